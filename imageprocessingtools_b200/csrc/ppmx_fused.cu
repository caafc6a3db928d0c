// ppmx_fused.cu -- one kernel for "geometry + pointwise tail": any of the eight orientations a chain of flips (ref:898-911)
// and 90/180/270 degree rotations (ref:714-725) composes to, followed by nothing, grey (ref:998-1000), the writer's .r
// extraction (ref:263-267) or mono with the P4 packer (ref:964-969 + 268-284) -- one read of the source, one write of
// what the writer emits, at ANY raster size and pointer alignment.  "ref:N" = /root/reference/ppmx-edward.c line N.
//
// A CTA moves one 64 x 64 pixel tile:
//   1. the tile's source rows come into shared memory as raw interleaved bytes -- by the bulk-copy engine
//      (cp.async.bulk + mbarrier) when rows are 16-byte aligned, else as aligned 16-byte vectors whose leading 0..15
//      bytes of misalignment are kept in the staged row and undone when pixels are cut out;
//   2. a thread takes 16 pixels -- down one source COLUMN for the four transposing orientations, along one source
//      ROW for the others --, applies the pointwise tail and writes them where they belong in an output tile in
//      shared memory (reversed when the orientation mirrors that axis);
//   3. the output tile's rows leave by the bulk-copy engine (16-byte aligned destination rows), as 8-byte vectors
//      (8-byte aligned, e.g. a 1080-pixel-wide result), or as aligned 16-byte vectors assembled by funnel shifts with
//      the up to 15 bytes at either end of a row piece written one by one (anything else).
// The Bayer index of mono, (x%4)*4 + (y%4) (ref:967), is taken in the coordinates of the stage where mono sits in
// the chain; the host passes that stage's coordinates as a signed permutation of the source coordinates.
#include "ppmx_common.cuh"

namespace ppmx {

constexpr int GI_PITCH = 208;  // staged source row: 192 B of pixels + up to 15 B of misalignment = 13 vectors, 52 words (rows spread over banks)
constexpr int GO_PITCH = 240;  // output tile row (RGB): 192 B + up to 45 B of front pad; 60 words: 16-byte stores of 8 rows never collide
constexpr int GO_PITCH_R8 = 80;
constexpr int GO_PITCH_BITS = 16;

enum { GP_RGB = 0, GP_GRAY = 1, GP_RED = 2, GP_MONO = 3 };
enum { GL_BULK = 0, GL_VEC = 1 };
enum { GS_BULK = 0, GS_VEC8 = 1, GS_SHIFT = 2 };

__device__ __forceinline__ uint32_t gs_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int POINT>
struct GeomOut {
    static constexpr int pitch = POINT == GP_RGB ? GO_PITCH : POINT == GP_MONO ? GO_PITCH_BITS : GO_PITCH_R8;
};

// A row piece of `nbytes` (<= 208) bytes from shared memory (4-byte aligned row, piece at byte `soff`) to any global
// address, by HALF a warp (hl = 0..15).
template <int STORE>
__device__ __forceinline__ void store_piece(uint8_t *g, const uint8_t *srow, uint32_t soff, uint32_t nbytes, uint32_t hl)
{
    if (STORE == GS_VEC8) {  // g, soff and nbytes are multiples of 8
        const uint2 *s = reinterpret_cast<const uint2 *>(srow + soff);
        for (uint32_t k = hl; k < nbytes / 8u; k += 16u) reinterpret_cast<uint2 *>(g)[k] = s[k];
        return;
    }
    // whole 4-byte words of the destination, coalesced over the half warp and each assembled from two shared-memory words by one
    // funnel shift; the up to 3 bytes in front of the first and behind the last word by lanes 0..2 and 3..5.  (The first form
    // wrote aligned 16-byte vectors and up to 15 bytes at either end one by one: 30 of a 192-byte piece went out as byte stores.)
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 3u);
    const uint32_t hb = min(nbytes, (4u - a) & 3u), nw = (nbytes - hb) >> 2, tb = nbytes - hb - 4u * nw;
    const uint32_t q0 = soff + hb, sh = 8u * (q0 & 3u);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(srow + (q0 & ~3u)) + hl;
    uint32_t *gw = reinterpret_cast<uint32_t *>(g + hb) + hl;
#pragma unroll
    for (uint32_t t = 0; t < 4u; t++)  // (nbytes <= 208: at most 52 words)
        if (hl + 16u * t < nw) gw[16u * t] = sh ? __funnelshift_r(sw[16u * t], sw[16u * t + 1], sh) : sw[16u * t];
    const bool tail = hl >= 3u;
    const uint32_t k = tail ? hl - 3u : hl, pos = tail ? hb + 4u * nw + k : k;
    if (hl < 6u && k < (tail ? tb : hb)) g[pos] = srow[soff + pos];
}

// TRANSPOSE 1: out row <- source column x (REV_Y: row w-1-x), out pixel <- source row y (REV_X: pixel h-1-y)
//           0: out row <- source row y    (REV_Y: row h-1-y), out pixel <- source column x (REV_X: pixel w-1-x)
template <int TRANSPOSE, bool REV_X, int POINT, int LOAD, int STORE>
__global__ void __launch_bounds__(128) geom_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t w,
                                                   uint32_t h, uint32_t in_pitch, uint32_t out_pitch, const GeomOp go)
{
    pdl_trigger();
    constexpr int OP = GeomOut<POINT>::pitch;
    __shared__ __align__(128) uint8_t tin[64 * GI_PITCH + 16];  // (+16: the funnel shift's second word of the last pixel)
    __shared__ __align__(128) uint8_t tout[64 * OP + 16];  // (+16: store_piece reads whole vectors)
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tx0 = blockIdx.x * 64u, ty0 = blockIdx.y * 64u;
    const uint32_t nc = min(64u, w - tx0), nr = min(64u, h - ty0);
    const uintptr_t g0 = reinterpret_cast<uintptr_t>(src) + (size_t)ty0 * in_pitch + (size_t)tx0 * 3;
    // (rows staged without the bulk-copy engine are byte-realigned on the way in: pixel tx0 of a row is byte 0 of its staged row)
    constexpr uint32_t a0 = 0u, ap = 0u;
    // ---- 1. the tile's source rows -> tin ------------------------------------------------------------------
    if (LOAD == GL_BULK) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs_smem(&bar)), "r"(1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        pdl_wait();
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gs_smem(&bar)), "r"(nr * nc * 3u) : "memory");
        const uint32_t crow = warp * 16u + (lane & 15u);  // a bulk copy is issued per warp: 16 per warp on all four warps
        if (lane < 16u && crow < nr)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             gs_smem(tin + crow * GI_PITCH)),
                         "l"(g0 + (size_t)crow * in_pitch), "r"(nc * 3u), "r"(gs_smem(&bar))
                         : "memory");
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(gs_smem(&bar)),
            "r"(0)
            : "memory");
    } else {
        pdl_wait();
        // Half a warp per row, four row pairs in flight.  A lane copies the words hl, hl+16, hl+32 of the row piece: each is cut
        // out of the two aligned global words that hold it (coalesced 4-byte loads, the second one an L1 hit of the neighbour's
        // first) by a funnel shift whose amount is the same for the whole row -- the misalignment is gone before the pixels
        // are cut out.  (The first form staged aligned 16-byte vectors and left every pixel read to undo its row's offset: ten
        // instructions per pixel instead of three.)
        const uint32_t hl = lane & 15u, hr = lane >> 4;
        const uint32_t nwords = (nc * 3u + 3u) >> 2;
        uint32_t *tin32w = reinterpret_cast<uint32_t *>(tin);
#pragma unroll
        for (uint32_t i4 = 0; i4 < 8u; i4 += 4u) {
            uint32_t lo[4][3], hi[4][3], sh[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t r = warp * 16u + 2u * (i4 + u) + hr;
                const uintptr_t g = g0 + (size_t)r * in_pitch;
                const uintptr_t last = (g + nc * 3u - 1u) & ~(uintptr_t)3;  // the last aligned word holding a byte of the piece
                const uint32_t *gw = reinterpret_cast<const uint32_t *>(g & ~(uintptr_t)3) + hl;
                sh[u] = 8u * (uint32_t)(g & 3u);
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const bool on = r < nr && hl + 16u * t < nwords;
                    lo[u][t] = on ? ld_global_u32(gw + 16 * t) : 0u;  // (not __ldg: see ld_global_u32)
                    // (a row that starts on a word boundary needs no second word: sh is the same for the whole half warp)
                    hi[u][t] = on && sh[u] != 0u && reinterpret_cast<uintptr_t>(gw + 16 * t + 1) <= last ? ld_global_u32(gw + 16 * t + 1) : 0u;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t r = warp * 16u + 2u * (i4 + u) + hr;
#pragma unroll
                for (int t = 0; t < 3; t++) tin32w[r * (GI_PITCH / 4) + hl + 16u * t] = __funnelshift_r(lo[u][t], hi[u][t], sh[u]);
            }
        }
        __syncthreads();
    }

    // ---- 2. 16 pixels per thread, pointwise tail, into the output tile --------------------------------------
    // the piece of an output row this tile holds starts `pad` pixels into the tout row when the pixel axis is mirrored
    // and the tile is ragged, so that the 16-pixel blocks of a thread never start before the row
    const uint32_t npx = TRANSPOSE ? nr : nc;  // pixels per output row piece
    const uint32_t pad = REV_X ? (16u - (npx & 15u)) & 15u : 0u;
    const uint32_t *tin32 = reinterpret_cast<const uint32_t *>(tin);
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        // TRANSPOSE: lane = source column, 16-row group j.   else: source row, 16-pixel group j.
        const uint32_t fix = TRANSPOSE ? lane + 32u * (warp & 1u) : (tid >> 2) + 32u * pass;
        const uint32_t j = TRANSPOSE ? (warp >> 1) + 2u * pass : (tid & 3u);
        if (TRANSPOSE ? (fix >= nc || 16u * j >= nr) : (fix >= nr || 16u * j >= nc)) continue;
        uint32_t px[16];  // r g b x per pixel (TRANSPOSE) ...
        uint32_t rw12[12];  // ... or 16 packed pixels (row-wise)
        if (TRANSPOSE) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t r = 16u * j + k, off = ((a0 + r * ap) & 15u) + 3u * fix;
                const uint32_t *p = tin32 + r * (GI_PITCH / 4) + (off >> 2);
                px[k] = __funnelshift_r(p[0], p[1], (off & 3u) * 8u);
            }
        } else {
            const uint32_t off = (a0 + fix * ap) & 15u;
            const uint32_t *p = tin32 + fix * (GI_PITCH / 4) + 12u * j + (off >> 2);
#pragma unroll
            for (int i = 0; i < 12; i++) rw12[i] = __funnelshift_r(p[i], p[i + 1], (off & 3u) * 8u);
        }
        // where the 16 pixels go: tout row `orow`, pixel block `oblk` of that row's piece (+ pad)
        const uint32_t orow = fix;
        const uint32_t oblk = REV_X ? (npx + pad) - 16u - 16u * j : 16u * j;
        if (POINT == GP_RGB) {
            uint32_t o[12];
            if (TRANSPOSE) {
#pragma unroll
                for (int i = 0; i < 4; i++) {  // 4 pixels -> 3 words
                    const uint32_t p0 = REV_X ? px[15 - 4 * i] : px[4 * i], p1 = REV_X ? px[14 - 4 * i] : px[4 * i + 1];
                    const uint32_t p2 = REV_X ? px[13 - 4 * i] : px[4 * i + 2], p3 = REV_X ? px[12 - 4 * i] : px[4 * i + 3];
                    o[3 * i] = __byte_perm(p0, p1, 0x4210);
                    o[3 * i + 1] = __byte_perm(p1, p2, 0x5421);
                    o[3 * i + 2] = __byte_perm(p2, p3, 0x6542);
                }
            } else if (REV_X) {
#pragma unroll
                for (int kk = 0; kk < 12; kk++) {  // reverse the order of 16 packed pixels
                    uint32_t v = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int ob = 4 * kk + b, ib = 3 * (15 - ob / 3) + (ob % 3);
                        v |= ((rw12[ib >> 2] >> (8 * (ib & 3))) & 0xFFu) << (8 * b);
                    }
                    o[kk] = v;
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 12; kk++) o[kk] = rw12[kk];
            }
            uint8_t *q = tout + orow * OP + oblk * 3u;  // 48-byte blocks: 16-byte aligned whenever pad * 3 is
            if (((pad * 3u) & 15u) == 0u) {
                uint4 *q4 = reinterpret_cast<uint4 *>(q);
                q4[0] = make_uint4(o[0], o[1], o[2], o[3]);
                q4[1] = make_uint4(o[4], o[5], o[6], o[7]);
                q4[2] = make_uint4(o[8], o[9], o[10], o[11]);
            } else if (((pad * 3u) & 3u) == 0u) {
#pragma unroll
                for (int kk = 0; kk < 12; kk++) reinterpret_cast<uint32_t *>(q)[kk] = o[kk];
            } else {
#pragma unroll
                for (int kk = 0; kk < 12; kk++)
#pragma unroll
                    for (int b = 0; b < 4; b++) q[4 * kk + b] = (uint8_t)(o[kk] >> (8 * b));
            }
        } else {
            // grey / red / Bayer bit of the 16 pixels, in SOURCE order along the thread's run
            uint32_t val[16];
            int thr3[4] = {0, 0, 0, 0};
            if (POINT == GP_MONO) {
                // run coordinate u = 16 j' + i has u % 4 == i % 4 (tiles start at multiples of 64); the other one is `fix`
                const uint32_t run0 = TRANSPOSE ? ty0 : tx0, fixc = (TRANSPOSE ? tx0 : ty0) + fix;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t sx = TRANSPOSE ? fixc : run0 + i, sy = TRANSPOSE ? run0 + i : fixc;  // source x, y (mod 4 is all that counts)
                    const uint32_t cx = go.mx_from_y ? sy : sx, cy = go.mx_from_y ? sx : sy;
                    const uint32_t xm = ((go.mx_neg ? 0u - cx : cx) + (uint32_t)go.mx_add) & 3u;
                    const uint32_t ym = ((go.my_neg ? 0u - cy : cy) + (uint32_t)go.my_add) & 3u;
                    thr3[i] = -3 * (int)c_bayer[xm * 4u + ym];
                }
            }
            if (TRANSPOSE) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (POINT == GP_GRAY) val[k] = div3(__dp4a(px[k], 0x00010101u, 0u));
                    else if (POINT == GP_RED) val[k] = px[k] & 0xFFu;
                    else val[k] = ((int)__dp4a(px[k], 0x00010101u, (uint32_t)thr3[k & 3]) < 0) ? 1u : 0u;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {  // 4 pixels in 3 words
                    const uint32_t a = rw12[3 * i], b = rw12[3 * i + 1], c = rw12[3 * i + 2];
                    const uint32_t t0 = POINT == GP_MONO ? (uint32_t)thr3[0] : 0u, t1 = POINT == GP_MONO ? (uint32_t)thr3[1] : 0u;
                    const uint32_t t2 = POINT == GP_MONO ? (uint32_t)thr3[2] : 0u, t3 = POINT == GP_MONO ? (uint32_t)thr3[3] : 0u;
                    if (POINT == GP_RED) {
                        val[4 * i] = a & 0xFFu;
                        val[4 * i + 1] = a >> 24;
                        val[4 * i + 2] = (b >> 16) & 0xFFu;
                        val[4 * i + 3] = (c >> 8) & 0xFFu;
                    } else {
                        const uint32_t s0 = __dp4a(a, 0x00010101u, t0), s1 = __dp4a(a, 0x01000000u, __dp4a(b, 0x00000101u, t1));
                        const uint32_t s2 = __dp4a(b, 0x01010000u, __dp4a(c, 0x00000001u, t2)), s3 = __dp4a(c, 0x01010100u, t3);
                        if (POINT == GP_GRAY) {
                            val[4 * i] = div3(s0);
                            val[4 * i + 1] = div3(s1);
                            val[4 * i + 2] = div3(s2);
                            val[4 * i + 3] = div3(s3);
                        } else {
                            val[4 * i] = s0 >> 31;
                            val[4 * i + 1] = s1 >> 31;
                            val[4 * i + 2] = s2 >> 31;
                            val[4 * i + 3] = s3 >> 31;
                        }
                    }
                }
            }
            if (POINT == GP_MONO) {
                uint32_t bits = 0;  // first OUTPUT pixel in the most significant bit (ref:273)
#pragma unroll
                for (int k = 0; k < 16; k++) bits |= val[REV_X ? 15 - k : k] << (15 - k);
                // pixels of this block beyond the raster's edge (a ragged last tile) are pad bits: zero (ref:268-284)
                const uint32_t valid = min(16u, npx - 16u * j);
                if (valid < 16u) bits &= REV_X ? (0xFFFFu >> (16u - valid)) : (0xFFFFu << (16u - valid));
                uint8_t *q = tout + orow * OP + (oblk >> 3);
                q[0] = (uint8_t)(bits >> 8);
                q[1] = (uint8_t)bits;
            } else {
                uint32_t o[4];
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const int i0 = REV_X ? 15 - 4 * kk : 4 * kk, st = REV_X ? -1 : 1;
                    o[kk] = val[i0] | (val[i0 + st] << 8) | (val[i0 + 2 * st] << 16) | (val[i0 + 3 * st] << 24);
                }
                uint8_t *q = tout + orow * OP + oblk;
                if ((pad & 15u) == 0u) *reinterpret_cast<uint4 *>(q) = make_uint4(o[0], o[1], o[2], o[3]);
                else if ((pad & 3u) == 0u) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) reinterpret_cast<uint32_t *>(q)[kk] = o[kk];
                } else {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
#pragma unroll
                        for (int b = 0; b < 4; b++) q[4 * kk + b] = (uint8_t)(o[kk] >> (8 * b));
                }
            }
        }
    }
    if (STORE == GS_BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the bulk engine must see the tile
    __syncthreads();

    // ---- 3. the output tile's rows -> dst -----------------------------------------------------------------------
    const uint32_t nrows_out = TRANSPOSE ? nc : nr;                       // rows of the output tile
    const uint32_t out_w = TRANSPOSE ? h : w, run0 = TRANSPOSE ? ty0 : tx0;
    const uint32_t xs = REV_X ? out_w - run0 - npx : run0;                // first output pixel of every row piece
    const uint32_t row0 = TRANSPOSE ? tx0 : ty0, out_h = TRANSPOSE ? w : h;
    uint32_t pbytes, soff;
    size_t goff;
    if (POINT == GP_RGB) {
        pbytes = npx * 3u, soff = pad * 3u, goff = (size_t)xs * 3;
    } else if (POINT == GP_MONO) {
        pbytes = (npx + 7u) >> 3, soff = pad >> 3, goff = xs >> 3;
    } else {
        pbytes = npx, soff = pad, goff = xs;
    }
    if (STORE == GS_BULK) {
        const uint32_t crow = warp * 16u + (lane & 15u);
        if (lane < 16u && crow < nrows_out) {
            const uint32_t y = go.rev_y ? out_h - 1u - (row0 + crow) : row0 + crow;
            uint8_t *g = dst + (size_t)y * out_pitch + goff;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(gs_smem(tout + crow * OP + soff)),
                         "r"(pbytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may go once it has been read
        }
    } else {
        const uint32_t hl = lane & 15u, hr = lane >> 4;  // half a warp per output row
        // the rows of one half warp are two apart: one 64-bit address, stepped by a constant
        const uint32_t crow0 = warp * 16u + hr;
        const uint32_t yfirst = go.rev_y ? out_h - 1u - (row0 + crow0) : row0 + crow0;
        uint8_t *g = dst + (size_t)yfirst * out_pitch + goff;
        const ptrdiff_t gstep = go.rev_y ? -2 * (ptrdiff_t)out_pitch : 2 * (ptrdiff_t)out_pitch;
        const uint8_t *srow = tout + crow0 * OP;
#pragma unroll 2
        for (uint32_t i = 0; i < 8u; i++, g += gstep, srow += 2 * OP) {
            if (crow0 + 2u * i >= nrows_out) break;
            store_piece<STORE>(g, srow, soff, pbytes, hl);
        }
    }
}

// ---- the four orientations that keep rows rows (identity, horizontal / vertical flip, 180 degrees) -----------------
// No tile is needed: a warp takes 32 consecutive 16-pixel groups of ONE source row (1536 bytes at any alignment).
//   load   a thread reads the four aligned 16-byte vectors that cover its 48 bytes; the row's misalignment m (0..15) is
//          the same for every thread of the row, so undoing it is a warp-uniform choice of a word offset (m / 4) and one
//          funnel-shift amount (m % 4);
//   tail   the same pointwise tails as the tile kernel (nothing, grey, .r, Bayer bits), pixel order reversed in
//          registers when the row is mirrored;
//   store  the warp's 32 results are one contiguous run of the destination row: they are laid down in a warp-private
//          piece of shared memory in destination order and leave as aligned 16-byte vectors assembled by funnel shifts
//          (up to 15 bytes at either end of the run byte by byte) -- rows of any pitch and alignment on both sides.
constexpr int RS_STAGE = 48 + 32 * 48 + 16;  // 16 pixels of front pad (a ragged mirrored group starts inside it) + the run + slack

template <int POINT, bool REV_X>
__global__ void __launch_bounds__(128) rows_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t w, uint32_t h,
                                                   uint32_t in_pitch, uint32_t out_pitch, const uint8_t *src_end, const GeomOp go)
{
    pdl_trigger();
    __shared__ __align__(16) uint8_t stage_all[4][RS_STAGE];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t y = blockIdx.y, ngroups = (w + 15u) >> 4;
    const uint32_t g0 = (blockIdx.x * 4u + warp) * 32u;  // the warp's first group
    if (g0 >= ngroups) return;
    const uint32_t grp = g0 + lane;
    uint8_t *stage = stage_all[warp];
    const uintptr_t A = reinterpret_cast<uintptr_t>(src) + (size_t)y * in_pitch + (size_t)grp * 48u;
    const uint32_t m = (uint32_t)((reinterpret_cast<uintptr_t>(src) + (size_t)y * in_pitch) & 15u);  // 48 * grp is a multiple of 16
    pdl_wait();

    // ---- load: four aligned vectors cover [A, A + 48) whatever m is; none starts at or beyond the raster's end ------
    uint32_t W[17];
    {
        const uint4 *vp = reinterpret_cast<const uint4 *>(A - m);
        const bool live = grp < ngroups;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (live && reinterpret_cast<const uint8_t *>(vp + k) < src_end && (k < 3 || m)) v = __ldg(vp + k);
            W[4 * k] = v.x, W[4 * k + 1] = v.y, W[4 * k + 2] = v.z, W[4 * k + 3] = v.w;
        }
        W[16] = 0u;  // (shift_words wants four spare words behind the twelve it returns: 12 + 3 + 1 <= 17)
    }
    uint32_t in[12];  // 16 pixels, packed
    shift_words<17, 12>(W, m, in);  // (m is the same for every group of the row)

    // ---- where this warp's run lies in the destination row -----------------------------------------------------------
    // pixels [16 g0, min(w, 16 g0 + 512)) of the source row; mirrored, the run starts at pixel max(0, w - 16 g0 - 512)
    // The staging area counts pixels from 16 before the run's first one (a front pad): the run starts at staging pixel 16.
    // Source pixel px0 + 16 lane + k goes to run pixel 16 lane + k, or mirrored to npx - 1 - 16 lane - k: the reversed block
    // of a group then starts at run pixel npx - 16 - 16 lane, which for the row's ragged last group lies inside the pad.
    const uint32_t px0 = 16u * g0, npx = min(512u, w - px0);
    const bool inrun = 16u * lane < npx;
    const uint32_t P = REV_X ? npx - 16u * lane : 16u + 16u * lane;  // this group's block, in staging pixels (inrun lanes only)
    const bool vec_ok = !REV_X || (npx & 15u) == 0u;                 // blocks sit on 16-byte boundaries of the staging area

    // ---- pointwise tail, in destination order ----------------------------------------------------------------------
    if (POINT == GP_RGB) {
        uint32_t o[12];
        if (REV_X) {
#pragma unroll
            for (int kk = 0; kk < 12; kk++) {  // reverse the order of 16 packed pixels: byte selections only (PRMT)
                const int ob0 = 4 * kk;
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int ob = ob0 + b, ib = 3 * (15 - ob / 3) + (ob % 3);
                    v |= ((in[ib >> 2] >> (8 * (ib & 3))) & 0xFFu) << (8 * b);
                }
                o[kk] = v;
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < 12; kk++) o[kk] = in[kk];
        }
        if (inrun) {
            uint8_t *q = stage + 3u * P;
            if (vec_ok) {
                uint4 *q4 = reinterpret_cast<uint4 *>(q);
                q4[0] = make_uint4(o[0], o[1], o[2], o[3]);
                q4[1] = make_uint4(o[4], o[5], o[6], o[7]);
                q4[2] = make_uint4(o[8], o[9], o[10], o[11]);
            } else {
#pragma unroll
                for (int kk = 0; kk < 12; kk++)
#pragma unroll
                    for (int b = 0; b < 4; b++) q[4 * kk + b] = (uint8_t)(o[kk] >> (8 * b));
            }
        }
    } else {
        uint32_t val[16];
        int thr3[4] = {0, 0, 0, 0};
        if (POINT == GP_MONO) {
#pragma unroll
            for (int i = 0; i < 4; i++) {  // x = 16 grp + 4 k + i: x % 4 == i
                const uint32_t sx = (uint32_t)i, sy = y;
                const uint32_t cx = go.mx_from_y ? sy : sx, cy = go.mx_from_y ? sx : sy;
                const uint32_t xm = ((go.mx_neg ? 0u - cx : cx) + (uint32_t)go.mx_add) & 3u;
                const uint32_t ym = ((go.my_neg ? 0u - cy : cy) + (uint32_t)go.my_add) & 3u;
                thr3[i] = -3 * (int)c_bayer[xm * 4u + ym];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {  // 4 pixels in 3 words
            const uint32_t a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
            if (POINT == GP_RED) {
                val[4 * i] = a & 0xFFu;
                val[4 * i + 1] = a >> 24;
                val[4 * i + 2] = (b >> 16) & 0xFFu;
                val[4 * i + 3] = (c >> 8) & 0xFFu;
            } else {
                const uint32_t t0 = POINT == GP_MONO ? (uint32_t)thr3[0] : 0u, t1 = POINT == GP_MONO ? (uint32_t)thr3[1] : 0u;
                const uint32_t t2 = POINT == GP_MONO ? (uint32_t)thr3[2] : 0u, t3 = POINT == GP_MONO ? (uint32_t)thr3[3] : 0u;
                const uint32_t s0 = __dp4a(a, 0x00010101u, t0), s1 = __dp4a(a, 0x01000000u, __dp4a(b, 0x00000101u, t1));
                const uint32_t s2 = __dp4a(b, 0x01010000u, __dp4a(c, 0x00000001u, t2)), s3 = __dp4a(c, 0x01010100u, t3);
                if (POINT == GP_GRAY) {
                    val[4 * i] = div3(s0), val[4 * i + 1] = div3(s1), val[4 * i + 2] = div3(s2), val[4 * i + 3] = div3(s3);
                } else {
                    val[4 * i] = s0 >> 31, val[4 * i + 1] = s1 >> 31, val[4 * i + 2] = s2 >> 31, val[4 * i + 3] = s3 >> 31;
                }
            }
        }
        if (POINT == GP_MONO) {
            uint32_t bits = 0;  // first DESTINATION pixel in the most significant bit (ref:273)
#pragma unroll
            for (int k = 0; k < 16; k++) bits |= val[REV_X ? 15 - k : k] << (15 - k);
            const uint32_t valid = 16u * lane < npx ? min(16u, npx - 16u * lane) : 0u;  // pixels beyond the row are pad bits: zero
            if (valid < 16u) bits &= REV_X ? (0xFFFFu >> (16u - valid)) : (0xFFFFu << (16u - valid));
            // Two bytes per lane, 64 per warp: straight to the destination row (no staging, no run store: for this tail they cost
            // more than the arithmetic).  Byte b of the run = staging byte b + 2 (a mirrored row of bits has w % 8 == 0, so P is a
            // multiple of 8; the ragged last group of a mirrored row starts one byte before the run).
            if (inrun) {
                const uint32_t oy = go.rev_y ? h - 1u - y : y, xs = REV_X ? w - px0 - npx : px0;
                const int b0 = (int)(P >> 3) - 2, nb = (int)((npx + 7u) >> 3);
                uint8_t *g = dst + (size_t)oy * out_pitch + (xs >> 3) + b0;
                if (b0 >= 0 && b0 < nb) g[0] = (uint8_t)(bits >> 8);
                if (b0 + 1 >= 0 && b0 + 1 < nb) g[1] = (uint8_t)bits;
            }
            return;
        } else {
            uint32_t o[4];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int i0 = REV_X ? 15 - 4 * kk : 4 * kk, st = REV_X ? -1 : 1;
                o[kk] = val[i0] | (val[i0 + st] << 8) | (val[i0 + 2 * st] << 16) | (val[i0 + 3 * st] << 24);
            }
            if (inrun) {
                uint8_t *q = stage + P;
                if (vec_ok) *reinterpret_cast<uint4 *>(q) = make_uint4(o[0], o[1], o[2], o[3]);
                else {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
#pragma unroll
                        for (int b = 0; b < 4; b++) q[4 * kk + b] = (uint8_t)(o[kk] >> (8 * b));
                }
            }
        }
    }
    __syncwarp();

    // ---- store: the run, to wherever the destination row puts it ----------------------------------------------------------
    const uint32_t oy = go.rev_y ? h - 1u - y : y;
    const uint32_t xs = REV_X ? w - px0 - npx : px0;  // first destination pixel of the run
    uint32_t nbytes, soff;
    size_t goff;
    if (POINT == GP_RGB) nbytes = npx * 3u, soff = 48u, goff = (size_t)xs * 3;
    else if (POINT == GP_MONO) nbytes = (npx + 7u) >> 3, soff = 2u, goff = xs >> 3;
    else nbytes = npx, soff = 16u, goff = xs;
    store_run(dst + (size_t)oy * out_pitch + goff, stage, soff, nbytes, lane);
}

template <int POINT, bool REV_X>
static cudaError_t rows_launch(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t in_pitch, uint32_t out_pitch,
                               const GeomOp &go, cudaStream_t s)
{
    const uint32_t ngroups = (w + 15u) / 16u;
    const uint8_t *src_end = src + (size_t)(h - 1) * in_pitch + (size_t)w * 3;
    for (uint32_t y0 = 0; y0 < h; y0 += 65535u) {  // rows ride on grid.y
        const uint32_t rows = h - y0 < 65535u ? h - y0 : 65535u;
        GeomOp g2 = go;
        // mono's Bayer phase counts whole-raster rows: shift it by the slab's first source row
        const int shift = (int)(y0 & 3u);
        if (g2.mx_from_y) g2.mx_add = (g2.mx_add + (g2.mx_neg ? 4 - shift : shift)) & 3;
        else g2.my_add = (g2.my_add + (g2.my_neg ? 4 - shift : shift)) & 3;
        // a vertically mirrored slab lands at the mirrored place
        uint8_t *d0 = dst + (size_t)(go.rev_y ? h - y0 - rows : y0) * out_pitch;
        dim3 grid((ngroups + 127u) / 128u, rows);
        launch(rows_kernel<POINT, REV_X>, grid, dim3(128), 0, s, src + (size_t)y0 * in_pitch, d0, w, rows, in_pitch, out_pitch, src_end, g2);
        cudaError_t e = PPMX_LAUNCHED();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int TRANSPOSE, bool REV_X, int POINT>
static cudaError_t geom_launch(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t in_pitch, uint32_t out_pitch,
                               const GeomOp &go, cudaStream_t s)
{
    const uint32_t out_w = TRANSPOSE ? h : w;
    const bool bulk_in = aligned16(src) && (in_pitch % 16u) == 0 && (w % 16u) == 0;
    int store = GS_SHIFT;
    if (POINT == GP_RGB) {
        if (aligned16(dst) && (out_pitch % 16u) == 0 && (out_w % 16u) == 0) store = GS_BULK;
        else if ((reinterpret_cast<uintptr_t>(dst) & 7u) == 0 && (out_pitch % 8u) == 0 && (out_w % 8u) == 0) store = GS_VEC8;
    } else if (POINT != GP_MONO && aligned16(dst) && (out_pitch % 16u) == 0 && (out_w % 16u) == 0) {
        store = GS_BULK;
    }
    dim3 grid((w + 63) / 64, (h + 63) / 64);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
#define PPMX_GEOM(L, S) launch(geom_kernel<TRANSPOSE, REV_X, POINT, L, S>, grid, dim3(128), 0, s, src, dst, w, h, in_pitch, out_pitch, go)
    if (bulk_in) {
        if (store == GS_BULK) PPMX_GEOM(GL_BULK, GS_BULK);
        else if (store == GS_VEC8) PPMX_GEOM(GL_BULK, GS_VEC8);
        else PPMX_GEOM(GL_BULK, GS_SHIFT);
    } else {
        if (store == GS_BULK) PPMX_GEOM(GL_VEC, GS_BULK);
        else if (store == GS_VEC8) PPMX_GEOM(GL_VEC, GS_VEC8);
        else PPMX_GEOM(GL_VEC, GS_SHIFT);
    }
#undef PPMX_GEOM
    return PPMX_LAUNCHED();
}

bool geom_point_supported(uint32_t w, uint32_t h, const GeomOp &go)
{
    if (go.point == GP_MONO) {
        // packed bits: a mirrored pixel axis must start on a byte boundary of the output row
        const uint32_t out_w = go.transpose ? h : w;
        if (go.rev_x && (out_w % 8u) != 0) return false;
    }
    return w > 0 && h > 0;
}

cudaError_t geom_point(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t in_pitch, const GeomOp &go, cudaStream_t s)
{
    if (!w || !h) return cudaSuccess;
    if (!geom_point_supported(w, h, go)) return cudaErrorInvalidValue;
    const uint32_t out_w = go.transpose ? h : w;
    const uint32_t out_pitch = go.point == GP_RGB ? out_w * 3u : go.point == GP_MONO ? (out_w + 7u) / 8u : out_w;
    if (!in_pitch) in_pitch = w * 3u;
#define PPMX_GEOM_P(T, R)                                                                          \
    switch (go.point) {                                                                            \
    case GP_RGB: return geom_launch<T, R, GP_RGB>(src, dst, w, h, in_pitch, out_pitch, go, s);     \
    case GP_GRAY: return geom_launch<T, R, GP_GRAY>(src, dst, w, h, in_pitch, out_pitch, go, s);   \
    case GP_RED: return geom_launch<T, R, GP_RED>(src, dst, w, h, in_pitch, out_pitch, go, s);     \
    case GP_MONO: return geom_launch<T, R, GP_MONO>(src, dst, w, h, in_pitch, out_pitch, go, s);   \
    default: return cudaErrorInvalidValue;                                                         \
    }
#define PPMX_ROWS_P(R)                                                                             \
    switch (go.point) {                                                                            \
    case GP_RGB: return rows_launch<GP_RGB, R>(src, dst, w, h, in_pitch, out_pitch, go, s);        \
    case GP_GRAY: return rows_launch<GP_GRAY, R>(src, dst, w, h, in_pitch, out_pitch, go, s);      \
    case GP_RED: return rows_launch<GP_RED, R>(src, dst, w, h, in_pitch, out_pitch, go, s);        \
    case GP_MONO: return rows_launch<GP_MONO, R>(src, dst, w, h, in_pitch, out_pitch, go, s);      \
    default: return cudaErrorInvalidValue;                                                         \
    }
    if (go.transpose) {
        if (go.rev_x) { PPMX_GEOM_P(1, true) } else { PPMX_GEOM_P(1, false) }
    } else {
        if (go.rev_x) { PPMX_ROWS_P(true) } else { PPMX_ROWS_P(false) }
    }
#undef PPMX_GEOM_P
#undef PPMX_ROWS_P
}

// rows of 3 * w bytes from one pitch to another (any alignment either side): the row kernel as a plain copy
cudaError_t geom_repitch(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t in_pitch, uint32_t out_pitch, cudaStream_t s)
{
    if (!w || !h) return cudaSuccess;
    GeomOp go = {};
    return rows_launch<GP_RGB, false>(src, dst, w, h, in_pitch, out_pitch, go, s);
}

}  // namespace ppmx
