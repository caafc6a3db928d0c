// ppmx_conv.cu -- EXTENSION (no reference counterpart, parity unpinned): k x k integer convolution with mirror border and row-band halos.
// Part of libppmx_gpu.so; see ppmx_common.cuh for conventions ("ref:N" = /root/reference/ppmx-edward.c line N).
#include "ppmx_conv.cuh"

namespace ppmx {

// ------------------------------------------------------------------------------------------
// EXTENSION (no reference counterpart, parity unpinned): k x k integer convolution.
// Border = symmetric mirror (the aux-table idiom of ref:551-555,589); result =
// floor(acc/div + 0.5) + bias computed in integers, clamped like ref:835.  Integer sums are
// exact in any order, so the fast kernel may regroup taps freely and still match the self-oracle.
// ------------------------------------------------------------------------------------------


// ---- generic kernel: any odd k <= 15, any coefficients, any width; scalar MACs ---------------
struct ConvCoefGeneric {
    int32_t c[CONV_MAXK * CONV_MAXK];
};
constexpr int CONV_TW = 64;  // output tile: 64 pixels x 16 rows per CTA
constexpr int CONV_TH = 16;

__global__ void __launch_bounds__(256) conv_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t w, int k,
                                                   int32_t div, int32_t bias, const ConvCoefGeneric cf)
{
    PDL_PROLOGUE();
    extern __shared__ uint8_t tile[];  // (CONV_TH + k - 1) rows x (CONV_TW + k - 1) pixels x 3
    const int r = k / 2, tw = CONV_TW + k - 1, th = CONV_TH + k - 1, tpitch = tw * 3;
    const int tx0 = blockIdx.x * CONV_TW, ty0 = blockIdx.y * CONV_TH;  // band-local output origin
    const size_t pitch = (size_t)w * 3;

    for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
        int ty = i / tw, tx = i - ty * tw;
        int gx = mirror_index(tx0 + tx - r, (int)w);
        const uint8_t *p = rs.row(rs.y0 + ty0 + ty - r, pitch) + (size_t)gx * 3;
        uint8_t *t = tile + ty * tpitch + tx * 3;
        t[0] = p[0];
        t[1] = p[1];
        t[2] = p[2];
    }
    __syncthreads();

    for (int i = threadIdx.x; i < CONV_TH * CONV_TW * 3; i += blockDim.x) {
        int ty = i / (CONV_TW * 3), b = i - ty * (CONV_TW * 3);
        int x = tx0 + b / 3, y = ty0 + ty;
        if (x >= (int)w || y >= rs.h) continue;
        long long acc = 0;
        for (int dy = 0; dy < k; dy++) {
            const uint8_t *t = tile + (ty + dy) * tpitch + b;
            for (int dx = 0; dx < k; dx++) acc += (long long)cf.c[dy * k + dx] * (int)t[dx * 3];
        }
        // floor((2*acc + div) / (2*div)) for div > 0
        long long num = 2 * acc + div, den = 2 * (long long)div, q = num / den;
        if ((num % den != 0) && (num < 0)) q--;
        q += bias;
        dst[(size_t)y * pitch + (size_t)tx0 * 3 + b] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
}

// ---- fast kernel: k in {3,5,7}, coefficients in [-128,127], w % 16 == 0, aligned rasters ------
// The tile is de-interleaved into three byte planes in shared memory, so horizontally adjacent taps
// of one channel are adjacent bytes and four of them feed one dp4a.  A thread produces 4 pixels x
// FC_RV rows x 3 channels.  Instead of shifting data to the tap window, the COEFFICIENTS are
// pre-shifted: for output j (0..3) of a 4-pixel word and source word wi (left, centre, right),
// cw[dy][j][wi] holds the 4 coefficients that multiply that word's bytes (0 where a tap does not
// reach).  Words that are all zero for a given (j, wi) are skipped at compile time.
constexpr int FC_TW = 128;  // tile width in pixels; the tile is 8 * RV rows tall (RV rows per thread)
constexpr int FC_PITCH = FC_TW + 32;  // one 16-pixel group of halo on each side


template <int K>
struct ConvCoefPacked {
    uint32_t cw[K][4][3];
    int32_t u[K];  // separable kernels (coef = u * v^T): cw[0] holds the words of v, u the column factor
};

template <int K>
__device__ __forceinline__ constexpr bool fc_reaches(int j, int wi)
{
    // does any byte b of source word wi (pixels 4(wi-1)+b relative to x) lie within K/2 of output j
    for (int b = 0; b < 4; b++) {
        int dx = 4 * (wi - 1) + b - j;
        if (dx >= -(K / 2) && dx <= K / 2) return true;
    }
    return false;
}

template <int K, int MODE, bool SEP, int FC_RV>
__global__ void __launch_bounds__(256) conv_dp4a_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t w,
                                                        const ConvCoefPacked<K> cf, const ConvRound rnd)
{
    PDL_PROLOGUE();
    constexpr int FC_TH = 8 * FC_RV, R = K / 2, IN_ROWS = FC_TH + K - 1;
    __shared__ __align__(16) uint8_t plane[3][IN_ROWS][FC_PITCH];
    const int tx0 = blockIdx.x * FC_TW, ty0 = blockIdx.y * FC_TH;
    const size_t pitch = (size_t)w * 3;

    // ---- stage: 16-pixel groups, IN_ROWS x (FC_PITCH / 16) of them ----
    constexpr int GROUPS_X = FC_PITCH / 16;
    // a tile that lies wholly inside the band and away from the raster's left/right edge (almost all of
    // them) needs no mirror, halo or bounds logic: its rows are consecutive rows of `own`
    const int gy_first = rs.y0 + ty0 - R;
    const bool inner = gy_first >= rs.y0 && gy_first + IN_ROWS <= rs.y0 + rs.h && tx0 >= 16 && tx0 + FC_TW + 16 <= (int)w;
    const uint8_t *inner_base = rs.own + (size_t)(gy_first - rs.y0) * pitch + (size_t)(tx0 - 16) * 3;
    for (int i = threadIdx.x; i < IN_ROWS * GROUPS_X; i += 256) {
        const int ty = i / GROUPS_X, gx = i - ty * GROUPS_X;
        const int x0 = tx0 - 16 + 16 * gx;  // first source pixel of the group (may be outside the raster)
        const uint8_t *row = inner ? inner_base + (size_t)ty * pitch - (size_t)(tx0 - 16) * 3
                                   : rs.row(gy_first + ty, pitch);
        if (inner || (x0 >= 0 && x0 + 16 <= (int)w)) {
            const uint4 *p = reinterpret_cast<const uint4 *>(row + (size_t)x0 * 3);
            const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
            const uint32_t wd[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
            uint32_t pr[4], pg[4], pb[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {  // 4 pixels (3 words) -> one word per plane
                const uint32_t u = wd[3 * q], v = wd[3 * q + 1], t = wd[3 * q + 2];
                pr[q] = __byte_perm(__byte_perm(u, v, 0x0630), t, 0x5210);  // u0 u3 v2 | t1
                pg[q] = __byte_perm(__byte_perm(u, v, 0x0741), t, 0x6210);  // u1 v0 v3 | t2
                pb[q] = __byte_perm(__byte_perm(u, v, 0x0052), t, 0x7410);  // u2 v1 | t0 t3
            }
            *reinterpret_cast<uint4 *>(&plane[0][ty][16 * gx]) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
            *reinterpret_cast<uint4 *>(&plane[1][ty][16 * gx]) = make_uint4(pg[0], pg[1], pg[2], pg[3]);
            *reinterpret_cast<uint4 *>(&plane[2][ty][16 * gx]) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
        } else {  // raster edge: mirrored columns, pixel by pixel
            for (int q = 0; q < 16; q++) {
                const uint8_t *p = row + (size_t)mirror_index(x0 + q, (int)w) * 3;
                plane[0][ty][16 * gx + q] = p[0];
                plane[1][ty][16 * gx + q] = p[1];
                plane[2][ty][16 * gx + q] = p[2];
            }
        }
    }
    __syncthreads();

    // ---- compute: thread = 4 pixels x FC_RV rows x 3 channels ----
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8 threads
    const int x = tx0 + 4 * lx, oy0 = ly * FC_RV;
    if (x >= (int)w) return;
    uint32_t outw[FC_RV][3];  // per output row: 4 result bytes of r, g, b
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        int32_t acc[FC_RV][4];
#pragma unroll
        for (int a = 0; a < FC_RV; a++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[a][j] = rnd.start;
#pragma unroll
        for (int iy = 0; iy < FC_RV + K - 1; iy++) {
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(&plane[ch][oy0 + iy][0]) + 3 + lx;
            const uint32_t wsrc[3] = {rw[0], rw[1], rw[2]};  // pixels x-4..x-1, x..x+3, x+4..x+7
            if (SEP) {
                // rank-1 kernel: one horizontal K-tap sum per source row (dp4a), then K cheap multiply-adds
                // spread it over the output rows it belongs to; the integer result is the same sum
                int32_t hsum[4] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int wi = 0; wi < 3; wi++)
                        if (fc_reaches<K>(j, wi)) hsum[j] = dp4a_u8s8(wsrc[wi], cf.cw[0][j][wi], hsum[j]);
#pragma unroll
                for (int a = 0; a < FC_RV; a++) {
                    const int dy = iy - a;
                    if (dy < 0 || dy >= K) continue;
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[a][j] += cf.u[dy] * hsum[j];
                }
            } else {
#pragma unroll
                for (int a = 0; a < FC_RV; a++) {
                    const int dy = iy - a;
                    if (dy < 0 || dy >= K) continue;
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int wi = 0; wi < 3; wi++)
                            if (fc_reaches<K>(j, wi)) acc[a][j] = dp4a_u8s8(wsrc[wi], cf.cw[dy][j][wi], acc[a][j]);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < FC_RV; a++)
            outw[a][ch] = rnd.template pack4<MODE>(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    }
#pragma unroll
    for (int a = 0; a < FC_RV; a++) {
        const int y = ty0 + oy0 + a;
        if (y >= rs.h) break;
        const uint32_t r4 = outw[a][0], g4 = outw[a][1], b4 = outw[a][2];
        // re-interleave 4 pixels: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
        const uint32_t o0 = __byte_perm(__byte_perm(r4, g4, 0x1040), b4, 0x3410);
        const uint32_t o1 = __byte_perm(__byte_perm(g4, b4, 0x0051), __byte_perm(r4, g4, 0x0062), 0x5410);
        const uint32_t o2 = __byte_perm(__byte_perm(b4, r4, 0x0072), __byte_perm(g4, b4, 0x0073), 0x5410);
        uint32_t *o = reinterpret_cast<uint32_t *>(dst + (size_t)y * pitch + (size_t)x * 3);
        o[0] = o0;
        o[1] = o1;
        o[2] = o2;
    }
}

// ---- 3x3 strip kernel: vertical dp4a on the interleaved raster, registers only ----------------
// The horizontal neighbour of a byte of an RGB raster is the byte 3 columns away, so no de-interleaving is needed
// if the dp4a runs DOWN the image instead of along it.  A thread owns 16 byte columns (one 16-byte vector per row)
// and walks down RH rows two at a time.  For every byte column it keeps the "vertical word"
//     V = in[y-1] | in[y] << 8 | in[y+1] << 16 | in[y+2] << 24          (rows around the output pair y, y+1)
// built from four source rows by 4x4 byte transposes (PRMT); successive pairs share two rows, so the first
// transpose stage of those is reused (6 PRMT per 4 columns per pair).  Output (y, c) is then three dp4a:
// V[c-3], V[c], V[c+3] against the coefficient column packed in bytes 0..2, output (y+1, c) the same words
// against the coefficients in bytes 1..3: 3 dp4a per output byte with 3 of the 4 multipliers busy (the
// row-wise form needs 4.5 and two (de)interleaving passes).  No shared memory, no barrier: the one word left
// and right of the thread's vector comes from L1 (it is the neighbouring lane's vector), or from the mirror
// rule at the raster's edge.  Instruction mix per output byte: 3 IDP (fma pipe), ~1.1 PRMT + 0.75..1.5
// finishing (alu pipe) -- both pipes issue 64 lanes/clk/SM (tools/int_peak.cu), so the kernel is bound by HBM.
// (struct Conv3Coef: ppmx_conv.cuh)

template <int MODE>
__device__ __forceinline__ uint32_t strip_pack4(const ConvRound &rnd, int32_t a0, int32_t a1, int32_t a2, int32_t a3)
{
    return rnd.template pack4<MODE>(a0, a1, a2, a3);
}

// UA ("unaligned"): rows of any length at any alignment (a width that is no multiple of 16, odd pointers).  A row's
// misalignment is the same for every thread of it, so a thread loads the aligned 16-byte vectors that cover its 24-byte
// window and undoes the shift with a warp-uniform word offset and one funnel-shift amount; the one or two chunks whose
// window crosses the row's end (and the first one) are read byte by byte with the mirror rule; the 16 result bytes of
// the warp's 32 threads are one run of the destination row, laid down in warp-private shared memory and written as
// aligned vectors (store_run) -- the padded copy the kernel used to need (two more passes over the raster) is gone.
constexpr int C3_STAGE = 32 * 16 + 16;

template <int MODE, int RH, int PF, bool INNER, bool WIDE, bool UA = false, bool BINOM = false>
__device__ __forceinline__ void conv3_strip_body(const RowSource &rs, uint8_t *__restrict__ dst, uint32_t nchunks, uint32_t cx,
                                                 int ys, const Conv3Coef &cf, const ConvRound &rnd, uint32_t row_bytes = 0,
                                                 uint8_t *stage = nullptr, uint32_t cx0 = 0)
{
    const size_t pitch = UA ? (size_t)row_bytes : (size_t)nchunks * 16;
    const bool left = cx == 0, right = cx == nchunks - 1;
    const int gy0 = rs.y0 + ys;
    const uint8_t *src = rs.own + (size_t)cx * 16 + (size_t)(INNER ? ys - 1 : 0) * pitch;
    // UA: the window [16 cx - 4, 16 cx + 20) of the first chunk starts before the row, that of the last ones ends beyond it
    const bool slow = UA && (cx == 0 || 16u * cx + 20u > row_bytes);
    auto load_row = [&](int i, uint32_t(&r)[6]) {  // row i counted from the strip's first source row (gy0 - 1)
        const uint8_t *p = INNER ? src + (size_t)i * pitch : rs.row(gy0 - 1 + i, pitch) + (size_t)cx * 16;
        if (UA) {
            if (slow) {
                const uint8_t *rowp = p - (size_t)cx * 16;
#pragma unroll
                for (int k = 0; k < 6; k++) r[k] = 0;
#pragma unroll
                for (int j = 0; j < 24; j++) {
                    int col = (int)(16u * cx) - 4 + j;
                    if (col < 0) col += 3;                         // pixel -1 mirrors to pixel 0
                    else if (col >= (int)row_bytes) col -= 3;      // pixel W to pixel W - 1
                    if (col < 0) col = 0;                          // (column -4: not used by any tap)
                    if (col >= (int)row_bytes) col = (int)row_bytes - 1;  // (bytes of this chunk beyond the row: never stored)
                    r[j >> 2] |= (uint32_t)rowp[col] << (8 * (j & 3));
                }
                return;
            }
            const uintptr_t a = reinterpret_cast<uintptr_t>(p) - 4u;
            const uint32_t o = (uint32_t)(a & 15u);
            const uint4 *vp = reinterpret_cast<const uint4 *>(a - o);
            const uint4 v0 = __ldg(vp), v1 = __ldg(vp + 1);
            uint4 v2 = make_uint4(0u, 0u, 0u, 0u);
            if (o > 8u) v2 = __ldg(vp + 2);  // 24 bytes from offset o end in the third vector
            const uint32_t W[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            shift_words<12, 6>(W, o, r);  // (o is the same for every thread of the row)
            return;
        }
        const uint4 m = __ldg(reinterpret_cast<const uint4 *>(p));
        r[1] = m.x, r[2] = m.y, r[3] = m.z, r[4] = m.w;
        // columns -3..-1 / 16..18; at the raster's edge pixel -1 mirrors to pixel 0 and pixel W to W-1
        r[0] = left ? (m.x << 8) : __ldg(reinterpret_cast<const uint32_t *>(p - 4));
        r[5] = right ? (m.w >> 8) : __ldg(reinterpret_cast<const uint32_t *>(p + 16));
    };
    pdl_wait();

    uint32_t tlo[6], thi[6];  // first transpose stage of the pair's two upper rows
    uint32_t nb[PF][2][6];    // the two new rows of the next PF pairs, in flight during the arithmetic
    {
        uint32_t a[6], b[6];
        load_row(0, a);
        load_row(1, b);
#pragma unroll
        for (int u = 0; u < PF; u++) {
            load_row(2 * u + 2, nb[u][0]);
            load_row(2 * u + 3, nb[u][1]);
        }
#pragma unroll
        for (int wc = 0; wc < 6; wc++) {
            tlo[wc] = __byte_perm(a[wc], b[wc], 0x5140);
            thi[wc] = __byte_perm(a[wc], b[wc], 0x7362);
        }
    }
    uint8_t *out = dst + (size_t)ys * pitch + (size_t)cx * 16;
#pragma unroll 1
    for (int g0 = 0; g0 < RH / 2; g0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int g = g0 + u;
            if (!INNER && ys + 2 * g >= rs.h) return;
            uint32_t ulo[6], uhi[6];
#pragma unroll
            for (int wc = 0; wc < 6; wc++) {
                ulo[wc] = __byte_perm(nb[u][0][wc], nb[u][1][wc], 0x5140);
                uhi[wc] = __byte_perm(nb[u][0][wc], nb[u][1][wc], 0x7362);
            }
            if (g + PF < RH / 2 && (INNER || ys + 2 * (g + PF) < rs.h)) {
                load_row(2 * (g + PF) + 2, nb[u][0]);
                load_row(2 * (g + PF) + 3, nb[u][1]);
            }
            uint32_t V[24];  // V[i] = byte column 16 cx - 4 + i
#pragma unroll
            for (int wc = 0; wc < 6; wc++) {
                V[4 * wc + 0] = __byte_perm(tlo[wc], ulo[wc], 0x5410);
                V[4 * wc + 1] = __byte_perm(tlo[wc], ulo[wc], 0x7632);
                V[4 * wc + 2] = __byte_perm(thi[wc], uhi[wc], 0x5410);
                V[4 * wc + 3] = __byte_perm(thi[wc], uhi[wc], 0x7632);
                tlo[wc] = ulo[wc];
                thi[wc] = uhi[wc];
            }
            uint32_t oa[4], ob[4];
            // BINOM (tuning build, variant 23; measured and NOT adopted): the filter is s * (1 2 1)^T (1 2 1).  One dot product per byte
            // column gives the vertical sum, the horizontal (1 2 1) is two adds: 1.4 dp4a + 2 adds per byte instead of 3 dp4a.  The idea
            // was that under the 1 kW cap fewer multipliers switching would keep the clock up; it is the instruction COUNT that
            // matters: 8192^2, 60 ms: 0.85 against 0.92; 2.6 s back to back: 0.76 against 0.82.
            int32_t SA[24], SB[24];
            if (BINOM) {
#pragma unroll
                for (int i = 1; i <= 22; i++) {
                    SA[i] = dp4a_u8s8(V[i], cf.a[0], 0);
                    SB[i] = dp4a_u8s8(V[i], cf.b[0], 0);
                }
            }
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int32_t accA[4], accB[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = 4 * b + j + 4;
                    if (BINOM) {
                        accA[j] = SA[c - 3] + SA[c + 3] + (2 * SA[c] + rnd.start);
                        accB[j] = SB[c - 3] + SB[c + 3] + (2 * SB[c] + rnd.start);
                        continue;
                    }
                    accA[j] = dp4a_u8s8(V[c + 3], cf.a[2], dp4a_u8s8(V[c], cf.a[1], dp4a_u8s8(V[c - 3], cf.a[0], rnd.start)));
                    accB[j] = dp4a_u8s8(V[c + 3], cf.b[2], dp4a_u8s8(V[c], cf.b[1], dp4a_u8s8(V[c - 3], cf.b[0], rnd.start)));
                    if (WIDE) {  // the high parts of the coefficients: a second dp4a chain, weighted 128
                        accA[j] += dp4a_u8s8(V[c + 3], cf.ah[2], dp4a_u8s8(V[c], cf.ah[1], dp4a_u8s8(V[c - 3], cf.ah[0], 0))) << 7;
                        accB[j] += dp4a_u8s8(V[c + 3], cf.bh[2], dp4a_u8s8(V[c], cf.bh[1], dp4a_u8s8(V[c - 3], cf.bh[0], 0))) << 7;
                    }
                }
                oa[b] = strip_pack4<MODE>(rnd, accA[0], accA[1], accA[2], accA[3]);
                ob[b] = strip_pack4<MODE>(rnd, accB[0], accB[1], accB[2], accB[3]);
            }
            if (UA) {
                // the warp's 32 x 16 result bytes = one run of the destination row, for each of the two rows
                const uint32_t lane = threadIdx.x & 31u, run = min(512u, row_bytes - 16u * cx0);
                const bool second = INNER || ys + 2 * g + 1 < rs.h;
                reinterpret_cast<uint4 *>(stage)[lane] = make_uint4(oa[0], oa[1], oa[2], oa[3]);
                reinterpret_cast<uint4 *>(stage + C3_STAGE)[lane] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
                __syncwarp();
                uint8_t *o0 = dst + (size_t)(ys + 2 * g) * pitch + (size_t)cx0 * 16;
                store_run(o0, stage, 0, run, lane);
                if (second) store_run(o0 + pitch, stage + C3_STAGE, 0, run, lane);
                __syncwarp();
            } else {
                *reinterpret_cast<uint4 *>(out) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
                if (INNER || ys + 2 * g + 1 < rs.h) *reinterpret_cast<uint4 *>(out + pitch) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
                out += 2 * pitch;
            }
        }
    }
}

template <int MODE, int RH, int PF, int BLOCK, bool WIDE = false, bool BINOM = false>
__global__ void __launch_bounds__(BLOCK) conv3_strip_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t nchunks,
                                                            const Conv3Coef cf, const ConvRound rnd)
{
    pdl_trigger();
    const uint32_t cx = blockIdx.x * BLOCK + threadIdx.x;
    if (cx >= nchunks) return;
    const int ys = blockIdx.y * RH;  // first output row of the strip, band-local
    // source rows ys-1 .. ys+RH all inside the own band (all strips but the first and last of a band): plain
    // pointer steps; otherwise every row goes through the mirror / halo resolver
    if (ys >= 1 && ys + RH + 1 <= rs.h) conv3_strip_body<MODE, RH, PF, true, WIDE, false, BINOM>(rs, dst, nchunks, cx, ys, cf, rnd);
    else conv3_strip_body<MODE, RH, PF, false, WIDE, false, BINOM>(rs, dst, nchunks, cx, ys, cf, rnd);
}

#ifdef PPMX_TUNING
// any width, any alignment: 4 rows per strip, one warp = 32 chunks of a row (lanes beyond the row's last chunk idle along)
template <int MODE, bool WIDE>
__global__ void __launch_bounds__(128) conv3_strip_ua_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t nchunks,
                                                             uint32_t row_bytes, const Conv3Coef cf, const ConvRound rnd)
{
    pdl_trigger();
    __shared__ __align__(16) uint8_t stage_all[4][2 * C3_STAGE];
    const uint32_t warp = threadIdx.x >> 5, cx0 = blockIdx.x * 128u + warp * 32u;
    if (cx0 >= nchunks) return;  // (a whole warp)
    const uint32_t cx = min(blockIdx.x * 128u + threadIdx.x, nchunks - 1u);
    const int ys = blockIdx.y * 4;
    if (ys >= 1 && ys + 4 + 1 <= rs.h) conv3_strip_body<MODE, 4, 2, true, WIDE, true>(rs, dst, nchunks, cx, ys, cf, rnd, row_bytes, stage_all[warp], cx0);
    else conv3_strip_body<MODE, 4, 2, false, WIDE, true>(rs, dst, nchunks, cx, ys, cf, rnd, row_bytes, stage_all[warp], cx0);
}
#endif

static cudaError_t conv3_strip(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef,
                               ConvRound rnd, int32_t div, int32_t bias, cudaStream_t s, bool unaligned = false)
{
    // A normalised non-negative filter with a power-of-two divisor (the blur presets) can not leave 0..255, so its
    // coefficients are pre-scaled by 256 / div and the quotient is byte 1 of the sum: three PRMT per four bytes instead
    // of a shift each plus two saturating packs.  Over a short run at full clocks the kernel is HBM-bound either way
    // (round 1 measured this form 2 % slower there); in a SUSTAINED run the 1 kW power cap holds the SMs at ~1.45-1.55
    // GHz, the alu pipe (60 % busy at full clock) becomes co-critical and every instruction saved shows.
    int32_t scaled_coef[9];
    int mode = rnd.mode;
    if (mode == 1 && bias == 0 && div >= 2 && div <= 256 && (256 % div) == 0 && PPMX_VARIANT != 8) {
        const int32_t k256 = 256 / div;
        int64_t sum = 0;
        bool ok = true;
        for (int i = 0; i < 9; i++) {
            ok = ok && coef[i] >= 0 && (int64_t)coef[i] * k256 <= 127;
            sum += coef[i];
        }
        if (ok && sum == div) {
            for (int i = 0; i < 9; i++) scaled_coef[i] = coef[i] * k256;
            coef = scaled_coef;
            rnd.start = 128;  // one half, in units of 1/256
            mode = 3;
        }
    }
    const int32_t *c = coef;
    Conv3Coef cf;
    bool wide = false;
    for (int dx = 0; dx < 3; dx++) {
        uint32_t lo[3], hi[3];
        for (int dy = 0; dy < 3; dy++) {  // c = 128 * hi + lo with lo in -64..63; hi == 0 for every byte-sized coefficient
            const int32_t v = c[3 * dy + dx];
            int32_t l = v, hpart = 0;
            if (v < -128 || v > 127) {
                l = ((v + 64) & 127) - 64;
                hpart = (v - l) / 128;
                wide = true;
            }
            lo[dy] = (uint32_t)(uint8_t)(int8_t)l;
            hi[dy] = (uint32_t)(uint8_t)(int8_t)hpart;
        }
        cf.a[dx] = lo[0] | lo[1] << 8 | lo[2] << 16;
        cf.b[dx] = lo[0] << 8 | lo[1] << 16 | lo[2] << 24;
        cf.ah[dx] = hi[0] | hi[1] << 8 | hi[2] << 16;
        cf.bh[dx] = hi[0] << 8 | hi[1] << 16 | hi[2] << 24;
    }
    if (unaligned && PPMX_VARIANT != 20) return conv3_ua_launch(rs, dst, w, h, cf, rnd, mode, wide, s);  // ppmx_conv_ua.cu
#ifdef PPMX_TUNING
    if (unaligned) {  // the first any-alignment kernel (three aligned vectors per thread and row, word selects): variant 20
        const uint32_t row_bytes = w * 3u, nch = (row_bytes + 15u) / 16u;
        dim3 grid((nch + 127) / 128, (h + 3) / 4);
        if (grid.y > 65535u) return cudaErrorInvalidValue;
#define PPMX_CONV3_UA(MODE)                                                                                             \
    do {                                                                                                                \
        if (wide) launch(conv3_strip_ua_kernel<MODE, true>, grid, dim3(128), 0, s, rs, dst, nch, row_bytes, cf, rnd);   \
        else launch(conv3_strip_ua_kernel<MODE, false>, grid, dim3(128), 0, s, rs, dst, nch, row_bytes, cf, rnd);       \
    } while (0)
        if (mode == 0) PPMX_CONV3_UA(0);
        else if (mode == 1) PPMX_CONV3_UA(1);
        else if (mode == 3) PPMX_CONV3_UA(3);
        else PPMX_CONV3_UA(2);
#undef PPMX_CONV3_UA
        return PPMX_LAUNCHED();
    }
#endif
    const uint32_t nchunks = w * 3 / 16;
#define PPMX_CONV3_LAUNCH(MODE, RH, PF, BLOCK)                                                                   \
    do {                                                                                                         \
        dim3 grid((nchunks + BLOCK - 1) / BLOCK, (h + RH - 1) / RH);                                             \
        if (grid.y > 65535u) return cudaErrorInvalidValue;                                                       \
        launch(conv3_strip_kernel<MODE, RH, PF, BLOCK>, grid, dim3(BLOCK), 0, s, rs, dst, nchunks, cf, rnd);     \
    } while (0)
#define PPMX_CONV3_MODES(RH, PF, BLOCK)                          \
    do {                                                         \
        if (mode == 0) PPMX_CONV3_LAUNCH(0, RH, PF, BLOCK);      \
        else if (mode == 1) PPMX_CONV3_LAUNCH(1, RH, PF, BLOCK); \
        else if (mode == 3) PPMX_CONV3_LAUNCH(3, RH, PF, BLOCK); \
        else PPMX_CONV3_LAUNCH(2, RH, PF, BLOCK);                \
    } while (0)
    // rows per strip / row pairs in flight / threads per CTA.  Measured on 8192x8192 (profiles/r1_sweep_conv3_strip.txt):
    // 4/2/128: 0.94 of the HBM roofline, 4/2/256: 0.92, 8/2/128: 0.84, 8/4/128: 0.84, 16/1/128: 0.74, 32/1/128: 0.68 --
    // the flatter the better (all six source rows of a strip are requested before the first dp4a, and the CTAs
    // resident at any moment cover a compact block of rows), as for the colour kernels.
    if (wide) {  // 6 instead of 3 dp4a per byte: still above what HBM delivers
        dim3 grid((nchunks + 127) / 128, (h + 3) / 4);
        if (grid.y > 65535u) return cudaErrorInvalidValue;
        if (mode == 0) launch(conv3_strip_kernel<0, 4, 2, 128, true>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, rnd);
        else if (mode == 1) launch(conv3_strip_kernel<1, 4, 2, 128, true>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, rnd);
        else launch(conv3_strip_kernel<2, 4, 2, 128, true>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, rnd);
        return PPMX_LAUNCHED();
    }
#ifdef PPMX_TUNING
    // s * (1 2 1)^T (1 2 1) in the byte-extraction form: vertical dot product + two adds (variant 23)
    if (PPMX_VARIANT == 23 && mode == 3 && cf.a[2] == cf.a[0] && cf.a[1] == 2u * cf.a[0] && cf.b[2] == cf.b[0] && cf.b[1] == 2u * cf.b[0]) {
        dim3 grid((nchunks + 127) / 128, (h + 3) / 4);
        if (grid.y > 65535u) return cudaErrorInvalidValue;
        launch(conv3_strip_kernel<3, 4, 2, 128, false, true>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, rnd);
        return PPMX_LAUNCHED();
    }
    if (PPMX_VARIANT == 9) PPMX_CONV3_MODES(2, 1, 128);
    else if (PPMX_VARIANT == 10) PPMX_CONV3_MODES(4, 2, 256);
    else if (PPMX_VARIANT == 11) PPMX_CONV3_MODES(8, 2, 128);
    else if (PPMX_VARIANT == 12) PPMX_CONV3_MODES(16, 1, 128);
    else if (PPMX_VARIANT == 13) PPMX_CONV3_MODES(4, 2, 64);
    else
#endif
    PPMX_CONV3_MODES(4, 2, 128);
#undef PPMX_CONV3_MODES
#undef PPMX_CONV3_LAUNCH
    return PPMX_LAUNCHED();
}

// ---- box kernel (all coefficients equal, e.g. the 7x7 blur preset): running sums ---------------
// A k x k box sum is a vertical running sum followed by a horizontal one, so its cost need not grow with k.
// A thread owns 16 byte columns and walks down RH rows keeping S[c] = sum of the K source rows around the
// current row for each of its columns; moving down one row is S += in[y+R+1] - in[y-R], one PRMT (pairing the
// leaving and the entering byte) and one dp4a against (-1, +1, 0, 0) per column.  The horizontal sum of S over
// the K pixels around a byte (columns 3 apart) needs 3R columns of the neighbouring threads: they come by warp
// shuffle, and so that no lane lacks a neighbour the warps overlap by two lanes (lanes 0 and 31 only supply
// their sums; 30 lanes write).  Along the row the sum slides too: h[c+3] = h[c] + S[c+3R+3] - S[c-3R], one
// IADD3.  The quotient is one multiply-add whose byte 3 is the result (constants verified on the host for every
// possible sum).  Source rows stream global -> shared with cp.async into thread-private slots of a 16-row ring
// (no barrier), requested 16-K rows ahead; the row that leaves the window is re-read from the ring, not from L2.
// Per output byte: ~2 fma-pipe (IDP, IMAD) + ~3.2 alu-pipe (PRMT, IADD3) + 1.1 SHFL, whatever K is.
constexpr int BOX_SLOTS = 16;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int K, int RH, bool INNER, bool EDGE>
__device__ __forceinline__ void conv_box_body(const RowSource &rs, uint8_t *__restrict__ dst, uint32_t nchunks, int chunk,
                                              int ys, uint32_t M, uint32_t C, uint4 (*ring)[128])
{
    constexpr int R = K / 2, H = 3 * R, TOTAL = RH + K - 1, PEND = BOX_SLOTS - K;
    constexpr int UNR = 4;  // rows per loop trip (a full unroll by the ring size made slot indices static but cost 90 KB of code)
    static_assert(RH % UNR == 0, "strip height");
    const int lane = threadIdx.x & 31;
    const bool valid = chunk >= 0 && chunk < (int)nchunks;
    const uint32_t cx = (uint32_t)(chunk < 0 ? 0 : chunk >= (int)nchunks ? (int)nchunks - 1 : chunk);
    const bool writes = valid && lane >= 1 && lane <= 30;
    const bool left = EDGE && chunk == 0, right = EDGE && chunk == (int)nchunks - 1;
    const size_t pitch = (size_t)nchunks * 16;
    const int gy0 = rs.y0 + ys;
    const uint8_t *src = rs.own + (size_t)cx * 16 + (size_t)(INNER ? ys - R : 0) * pitch;
    auto row_ptr = [&](int i) {  // row i counted from the strip's first source row (gy0 - R)
        return INNER ? src + (size_t)i * pitch : rs.row(gy0 - R + i, pitch) + (size_t)cx * 16;
    };
    uint4 *slot0 = &ring[0][threadIdx.x];
    auto slot = [&](int i) { return slot0 + (size_t)(i & (BOX_SLOTS - 1)) * 128; };
    // rows a strip shortened by the band's end never uses are still fetched (the resolver mirrors them): harmless
    pdl_wait();
#pragma unroll
    for (int i = 0; i < BOX_SLOTS; i++) {
        if (i < TOTAL) cp_async16(slot(i), row_ptr(i));
        cp_async_commit();
    }
    cp_async_wait<PEND>();

    // S = sum of rows 0..K-1, four rows at a time through 4x4 byte transposes and dp4a against (1,1,1,1)
    uint32_t S[16];
#pragma unroll
    for (int c = 0; c < 16; c++) S[c] = 0;
#pragma unroll
    for (int i0 = 0; i0 < K; i0 += 4) {
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; u++) q[u] = i0 + u < K ? *slot(i0 + u) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int wc = 0; wc < 4; wc++) {
            const uint32_t a = (&q[0].x)[wc], b = (&q[1].x)[wc], c = (&q[2].x)[wc], d = (&q[3].x)[wc];
            const uint32_t t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(c, d, 0x5140);
            const uint32_t t2 = __byte_perm(a, b, 0x7362), t3 = __byte_perm(c, d, 0x7362);
            S[4 * wc + 0] = __dp4a(__byte_perm(t0, t1, 0x5410), 0x01010101u, S[4 * wc + 0]);
            S[4 * wc + 1] = __dp4a(__byte_perm(t0, t1, 0x7632), 0x01010101u, S[4 * wc + 1]);
            S[4 * wc + 2] = __dp4a(__byte_perm(t2, t3, 0x5410), 0x01010101u, S[4 * wc + 2]);
            S[4 * wc + 3] = __dp4a(__byte_perm(t2, t3, 0x7632), 0x01010101u, S[4 * wc + 3]);
        }
    }

    uint8_t *out = dst + (size_t)ys * pitch + (size_t)cx * 16;
#pragma unroll 1
    for (int r0 = 0; r0 < RH; r0 += UNR) {
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int r = r0 + u;
            if (!INNER && ys + r >= rs.h) return;
            // ---- the slot of the row that left the window one step ago takes the next row to fetch ----
            if (r0 + u >= 1) {  // (one group per step, empty or not, keeps the wait count below a constant)
                if (BOX_SLOTS + r - 1 < TOTAL) cp_async16(slot(r - 1), row_ptr(BOX_SLOTS + r - 1));
                cp_async_commit();
            }
            // ---- horizontal sums: 3R columns from the lane on either side ----
            uint32_t XL[H], XR[H];  // S of byte columns -H .. -1 and 16 .. 15 + H
#pragma unroll
            for (int i = 0; i < H; i++) {
                XL[i] = __shfl_up_sync(0xffffffffu, S[16 - H + i], 1);
                XR[i] = __shfl_down_sync(0xffffffffu, S[i], 1);
            }
            if (EDGE) {  // only the warps holding the raster's first or last 16 bytes compile this in
#pragma unroll
                for (int i = 0; i < H; i++) {
                    // pixel -m mirrors to pixel m-1: columns -3m+j come from columns 3(m-1)+j;
                    // pixel W-1+m mirrors to pixel W-m: columns 16+i come from 13 - 3(i/3) + i%3
                    XL[i] = left ? S[3 * (R - 1 - i / 3) + i % 3] : XL[i];
                    XR[i] = right ? S[13 - 3 * (i / 3) + i % 3] : XR[i];
                }
            }
            auto X = [&](int c) { return c < 0 ? XL[c + H] : c < 16 ? S[c] : XR[c - 16]; };
            uint32_t hs[16];
#pragma unroll
            for (int c = 0; c < 16; c++) {
                if (c < 3) {
                    uint32_t t = 0;
#pragma unroll
                    for (int k = -R; k <= R; k++) t += X(c + 3 * k);
                    hs[c] = t;
                } else {
                    hs[c] = hs[c - 3] + X(c + 3 * R) - X(c - 3 * R - 3);
                }
            }
            uint32_t o[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t q0 = hs[4 * b] * M + C, q1 = hs[4 * b + 1] * M + C, q2 = hs[4 * b + 2] * M + C,
                               q3 = hs[4 * b + 3] * M + C;  // the quotient is byte 3
                o[b] = __byte_perm(__byte_perm(q0, q1, 0x0073), __byte_perm(q2, q3, 0x0073), 0x5410);
            }
            if (writes) *reinterpret_cast<uint4 *>(out) = make_uint4(o[0], o[1], o[2], o[3]);
            out += pitch;
            if (r == RH - 1) return;
            // ---- move the window down: S += row (K + r) - row r, one byte-selecting dp4a each ----
            cp_async_wait<PEND - 1>();  // groups 0 .. K + r have landed
            const uint4 nw = *slot(K + r), od = *slot(r);
#pragma unroll
            for (int wc = 0; wc < 4; wc++) {
                const uint32_t a = (&od.x)[wc], b = (&nw.x)[wc];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    S[4 * wc + j] = dp4a_u8s8(b, 0x01u << (8 * j), dp4a_u8s8(a, 0xffu << (8 * j), S[4 * wc + j]));
            }
        }
    }
}

template <int K, int RH>
__global__ void __launch_bounds__(128, 7) conv_box_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t nchunks, uint32_t M,
                                                       uint32_t C)
{
    constexpr int R = K / 2;
    __shared__ uint4 ring[BOX_SLOTS][128];
    pdl_trigger();
    const int wg = blockIdx.x * 4 + (threadIdx.x >> 5);  // warp number along the row: 30 chunks each
    if (wg * 30 >= (int)nchunks) return;
    const int chunk = wg * 30 - 1 + (int)(threadIdx.x & 31);
    const int ys = blockIdx.y * RH;
    const bool inner = ys >= R && ys + RH + R <= rs.h;
    const bool edge = wg == 0 || wg * 30 + 30 >= (int)nchunks;  // this warp holds the raster's first or last chunk
    if (inner && !edge) conv_box_body<K, RH, true, false>(rs, dst, nchunks, chunk, ys, M, C, ring);
    else if (inner) conv_box_body<K, RH, true, true>(rs, dst, nchunks, chunk, ys, M, C, ring);
    else conv_box_body<K, RH, false, true>(rs, dst, nchunks, chunk, ys, M, C, ring);
}

// all coefficients equal to a > 0, and (sum * M + C) >> 24 == floor((2 a sum + div) / (2 div)) + bias within 0..255
// for every possible sum (checked one by one: at most 255 k^2 + 1 values)
static bool box_constants(const int32_t *coef, int k, int32_t div, int32_t bias, uint32_t *M, uint32_t *C)
{
    const int32_t a = coef[0];
    if (a <= 0) return false;
    for (int i = 1; i < k * k; i++)
        if (coef[i] != a) return false;
    const int64_t smax = 255ll * k * k;
    const uint64_t m = (((uint64_t)a << 24) + (uint64_t)div - 1) / (uint64_t)div;  // ceil(2^24 a / div)
    // floor((2 a s + div) / (2 div)) = floor((a s + div / 2) / div) for integers, div / 2 rounded down
    const int64_t c = (int64_t)((((uint64_t)(div / 2) << 24) + (uint64_t)div - 1) / (uint64_t)div) + ((int64_t)bias << 24);
    if (c < 0 || m >= (1ull << 32) || (uint64_t)smax * m + (uint64_t)c >= (1ull << 32)) return false;
    for (int64_t sum = 0; sum <= smax; sum++) {
        const int64_t num = 2 * (int64_t)a * sum + div, den = 2 * (int64_t)div;
        const int64_t q = num / den + bias;  // num >= 0
        if (q < 0 || q > 255) return false;
        if ((int64_t)(((uint64_t)sum * m + (uint64_t)c) >> 24) != q) return false;
    }
    *M = (uint32_t)m;
    *C = (uint32_t)c;
    return true;
}

template <int K>
static cudaError_t conv_box(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, uint32_t M, uint32_t C, cudaStream_t s)
{
    const uint32_t nchunks = w * 3 / 16;
#define PPMX_BOX_LAUNCH(RH)                                                                   \
    do {                                                                                      \
        dim3 grid((nchunks + 119) / 120, (h + RH - 1) / RH);                                  \
        if (grid.y > 65535u) return cudaErrorInvalidValue;                                    \
        launch(conv_box_kernel<K, RH>, grid, dim3(128), 0, s, rs, dst, nchunks, M, C);        \
    } while (0)
#ifdef PPMX_TUNING
    if (PPMX_VARIANT == 10) PPMX_BOX_LAUNCH(32);
    else if (PPMX_VARIANT == 11) PPMX_BOX_LAUNCH(64);
    else
#endif
    PPMX_BOX_LAUNCH(16);
#undef PPMX_BOX_LAUNCH
    return PPMX_LAUNCHED();
}


// ---- rank-1 kernels (coef = u v^T: Gaussian / binomial blurs, Sobel-like derivatives), K = 5 or 7 ----------
// Same thread layout as the box kernel (16 byte columns per thread walking down RH rows, warps overlapped by two
// lanes, neighbours' column sums by shuffle), but the vertical pass is a real dot product.  Per byte column the
// thread keeps the last 8 source rows as two sliding "vertical words" (W_lo = rows y+R-7 .. y+R-4, W_hi = rows
// y+R-3 .. y+R): a new row shifts them with two PRMT, and the K-tap column sum is two dp4a against u packed into
// bytes -- no shared memory, no re-reads.  The horizontal pass multiplies the 32-bit column sums by v, with the
// mirror pairs added first when v is symmetric (R adds + R+1 IMAD instead of K IMAD).
// Per output byte at K = 7: 2 IDP + 4 IMAD on the fma pipe, 2 PRMT + 3 IADD3 + finishing on the alu pipe, 1.1 SHFL.
template <int K>
struct SepCoef {
    uint32_t u_lo, u_hi;  // u[dy] in the byte of the row it multiplies: W_lo byte b <-> dy = R-7+b, W_hi byte b <-> dy = R-3+b
    int32_t v[K];
};

template <int K, int MODE, bool SYM, int RH, bool INNER, bool EDGE>
__device__ __forceinline__ void conv_sep_body(const RowSource &rs, uint8_t *__restrict__ dst, uint32_t nchunks, int chunk, int ys,
                                              const SepCoef<K> &cf, const ConvRound &rnd)
{
    constexpr int R = K / 2, H = 3 * R, PF = 2;
    static_assert(RH % PF == 0, "strip height");
    const int lane = threadIdx.x & 31;
    const bool valid = chunk >= 0 && chunk < (int)nchunks;
    const uint32_t cx = (uint32_t)(chunk < 0 ? 0 : chunk >= (int)nchunks ? (int)nchunks - 1 : chunk);
    const bool writes = valid && lane >= 1 && lane <= 30;
    const bool left = EDGE && chunk == 0, right = EDGE && chunk == (int)nchunks - 1;
    const size_t pitch = (size_t)nchunks * 16;
    const int gy0 = rs.y0 + ys;
    const uint8_t *src = rs.own + (size_t)cx * 16 + (size_t)(INNER ? ys - R : 0) * pitch;
    auto load_row = [&](int i) {  // row i counted from the strip's first source row (gy0 - R)
        const uint8_t *p = INNER ? src + (size_t)i * pitch : rs.row(gy0 - R + i, pitch) + (size_t)cx * 16;
        return __ldg(reinterpret_cast<const uint4 *>(p));
    };
    uint32_t wlo[16], whi[16];
#pragma unroll
    for (int c = 0; c < 16; c++) wlo[c] = whi[c] = 0;
    auto push = [&](const uint4 &row) {  // the window moves down one row
        const uint32_t rw[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
        for (int c = 0; c < 16; c++) {
            if (K > 4) wlo[c] = __byte_perm(wlo[c], whi[c], 0x4321);
            whi[c] = __byte_perm(whi[c], rw[c >> 2], 0x0321 | ((4 + (c & 3)) << 12));
        }
    };
    pdl_wait();
    {
        uint4 first[2 * R];
#pragma unroll
        for (int i = 0; i < 2 * R; i++) first[i] = load_row(i);
#pragma unroll
        for (int i = 0; i < 2 * R; i++) push(first[i]);
    }
    uint4 nx[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) nx[u] = load_row(2 * R + u);

    uint8_t *out = dst + (size_t)ys * pitch + (size_t)cx * 16;
#pragma unroll 1
    for (int r0 = 0; r0 < RH; r0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int r = r0 + u;
            if (!INNER && ys + r >= rs.h) return;
            push(nx[u]);
            if (r + PF < RH && (INNER || ys + r + PF < rs.h)) nx[u] = load_row(2 * R + r + PF);
            int32_t S[16];
#pragma unroll
            for (int c = 0; c < 16; c++) S[c] = dp4a_u8s8(whi[c], cf.u_hi, K > 4 ? dp4a_u8s8(wlo[c], cf.u_lo, 0) : 0);
            int32_t XL[H], XR[H];  // column sums of byte columns -H .. -1 and 16 .. 15 + H
#pragma unroll
            for (int i = 0; i < H; i++) {
                XL[i] = __shfl_up_sync(0xffffffffu, S[16 - H + i], 1);
                XR[i] = __shfl_down_sync(0xffffffffu, S[i], 1);
            }
            if (EDGE) {  // mirror at the raster's left / right edge, as in the box kernel
#pragma unroll
                for (int i = 0; i < H; i++) {
                    XL[i] = left ? S[3 * (R - 1 - i / 3) + i % 3] : XL[i];
                    XR[i] = right ? S[13 - 3 * (i / 3) + i % 3] : XR[i];
                }
            }
            auto X = [&](int c) { return c < 0 ? XL[c + H] : c < 16 ? S[c] : XR[c - 16]; };
            uint32_t o[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int32_t a[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = 4 * b + j;
                    int32_t t = rnd.start + cf.v[R] * X(c);
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        if (SYM) t += cf.v[R + k] * (X(c - 3 * k) + X(c + 3 * k));
                        else t += cf.v[R - k] * X(c - 3 * k) + cf.v[R + k] * X(c + 3 * k);
                    }
                    a[j] = t;
                }
                o[b] = rnd.template pack4<MODE>(a[0], a[1], a[2], a[3]);
            }
            if (writes) *reinterpret_cast<uint4 *>(out) = make_uint4(o[0], o[1], o[2], o[3]);
            out += pitch;
        }
    }
}

template <int K, int MODE, bool SYM, int RH>
__global__ void __launch_bounds__(128, 4) conv_sep_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t nchunks, const SepCoef<K> cf,
                                                       const ConvRound rnd)
{
    constexpr int R = K / 2;
    pdl_trigger();
    const int wg = blockIdx.x * 4 + (threadIdx.x >> 5);  // warp number along the row: 30 chunks each
    if (wg * 30 >= (int)nchunks) return;
    const int chunk = wg * 30 - 1 + (int)(threadIdx.x & 31);
    const int ys = blockIdx.y * RH;
    const bool inner = ys >= R && ys + RH + R <= rs.h;
    const bool edge = wg == 0 || wg * 30 + 30 >= (int)nchunks;
    if (inner && !edge) conv_sep_body<K, MODE, SYM, RH, true, false>(rs, dst, nchunks, chunk, ys, cf, rnd);
    else if (inner) conv_sep_body<K, MODE, SYM, RH, true, true>(rs, dst, nchunks, chunk, ys, cf, rnd);
    else conv_sep_body<K, MODE, SYM, RH, false, true>(rs, dst, nchunks, chunk, ys, cf, rnd);
}

template <int K>
static bool rank_one(const int32_t *coef, int32_t (&u)[K], int32_t (&v)[K]);

// returns false when the kernel is not rank 1 with a column factor that fits signed bytes
template <int K, int RH>
static bool conv_sep_rh(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef, const ConvRound &rnd,
                     cudaStream_t s, cudaError_t *err)
{
    constexpr int R = K / 2;
    int32_t u[K], v[K];
    if (!rank_one<K>(coef, u, v)) return false;
    SepCoef<K> cf;
    cf.u_lo = cf.u_hi = 0;
    bool sym = true;
    for (int i = 0; i < K; i++) {
        if (u[i] < -128 || u[i] > 127) return false;
        const int dy = i - R, bh = dy - (R - 3), bl = dy - (R - 7);
        if (bh >= 0 && bh < 4) cf.u_hi |= (uint32_t)(uint8_t)(int8_t)u[i] << (8 * bh);
        else if (bl >= 0 && bl < 4) cf.u_lo |= (uint32_t)(uint8_t)(int8_t)u[i] << (8 * bl);
        else return false;
        cf.v[i] = v[i];
        sym = sym && v[i] == v[K - 1 - i];
    }
    const uint32_t nchunks = w * 3 / 16;
    dim3 grid((nchunks + 119) / 120, (h + RH - 1) / RH);
    if (grid.y > 65535u) {
        *err = cudaErrorInvalidValue;
        return true;
    }
#define PPMX_SEP_LAUNCH(MODE, SYM) launch(conv_sep_kernel<K, MODE, SYM, RH>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, rnd)
    if (sym) {
        if (rnd.mode == 0) PPMX_SEP_LAUNCH(0, true);
        else if (rnd.mode == 1) PPMX_SEP_LAUNCH(1, true);
        else PPMX_SEP_LAUNCH(2, true);
    } else {
        if (rnd.mode == 0) PPMX_SEP_LAUNCH(0, false);
        else if (rnd.mode == 1) PPMX_SEP_LAUNCH(1, false);
        else PPMX_SEP_LAUNCH(2, false);
    }
#undef PPMX_SEP_LAUNCH
    *err = PPMX_LAUNCHED();
    return true;
}

template <int K>
static bool conv_sep(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef, const ConvRound &rnd,
                     cudaStream_t s, cudaError_t *err)
{
    // rows per strip, 7x7 / 5x5 binomial at 8192^2: 16 -> 0.46 / 0.53 of the HBM roofline, 32 -> 0.45 / 0.52, 64 -> 0.42 / 0.48
#ifdef PPMX_TUNING
    if (PPMX_VARIANT == 9) return conv_sep_rh<K, 8>(rs, dst, w, h, coef, rnd, s, err);
    if (PPMX_VARIANT == 10) return conv_sep_rh<K, 32>(rs, dst, w, h, coef, rnd, s, err);
#endif
    return conv_sep_rh<K, 16>(rs, dst, w, h, coef, rnd, s, err);
}

template <int K>
static cudaError_t conv_fast(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef,
                             const ConvRound &rnd, cudaStream_t s)
{
    ConvCoefPacked<K> cf;
    int32_t u[K], v[K];
    // 3x3: the direct form measured faster (0.51 vs 0.49 of the HBM roofline); 5x5 and 7x7: separable wins
    const bool sep = PPMX_VARIANT != 2 && K >= 5 && rank_one<K>(coef, u, v);
    for (int dy = 0; dy < K; dy++) {
        cf.u[dy] = sep ? u[dy] : 0;
        for (int j = 0; j < 4; j++)
            for (int wi = 0; wi < 3; wi++) {
                uint32_t word = 0;
                for (int b = 0; b < 4; b++) {
                    int dx = 4 * (wi - 1) + b - j;
                    if (dx >= -(K / 2) && dx <= K / 2) {
                        const int32_t c = sep ? v[dx + K / 2] : coef[dy * K + dx + K / 2];
                        word |= (uint32_t)(uint8_t)(int8_t)c << (8 * b);
                    }
                }
                cf.cw[dy][j][wi] = word;
            }
    }
    // separable kernels amortise the horizontal sums over more rows per thread (8 instead of 4)
    constexpr int RV_SEP = 8, RV_DIR = 4;  // (8 rows per thread for the direct form measured slower: 0.47 vs 0.51)
    const int th = 8 * (sep ? RV_SEP : RV_DIR);
    dim3 grid((w + FC_TW - 1) / FC_TW, (h + th - 1) / th);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
#define PPMX_CONV_LAUNCH(MODE)                                                                                      \
    do {                                                                                                            \
        if (sep) launch(conv_dp4a_kernel<K, MODE, true, RV_SEP>, grid, dim3(256), 0, s, rs, dst, w, cf, rnd);      \
        else launch(conv_dp4a_kernel<K, MODE, false, RV_DIR>, grid, dim3(256), 0, s, rs, dst, w, cf, rnd);         \
    } while (0)
    if (rnd.mode == 0) PPMX_CONV_LAUNCH(0);
    else if (rnd.mode == 1) PPMX_CONV_LAUNCH(1);
    else PPMX_CONV_LAUNCH(2);
#undef PPMX_CONV_LAUNCH
    return PPMX_LAUNCHED();
}

// ---- rasters whose width or pointers do not allow 16-byte vectors (w % 16 != 0, odd addresses) ----------------
// The vector kernels want a row pitch that is a multiple of 16 bytes.  Such a raster is copied into a pool buffer
// of width wp = roundup(w + k/2, 16) with the columns beyond the right edge filled by the mirror rule (so every real
// pixel sees exactly the neighbours the border rule gives it), convolved there by the fast kernels, and the real
// columns are copied out again: two extra passes through the copy engine's 2-D path.
// The rows themselves move by cudaMemcpy2DAsync (device to device, any pitch); this kernel only fills the
// wp - w mirrored pixels at the end of every padded row.
__global__ void __launch_bounds__(96) conv_pad_edge_kernel(uint8_t *__restrict__ padded, uint32_t w, uint32_t wp)
{
    PDL_PROLOGUE();
    const uint32_t t = threadIdx.x, y = blockIdx.x;  // byte t of the mirrored tail of padded row y
    if (t >= (wp - w) * 3u) return;
    uint8_t *row = padded + (size_t)y * wp * 3;
    const uint32_t px = w + t / 3u, ch = t % 3u;
    row[px * 3u + ch] = row[(uint32_t)mirror_index((int)px, (int)w) * 3u + ch];
}

cudaError_t conv(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div,
                 int32_t bias, const Band &band, cudaStream_t s);

static cudaError_t conv_padded(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div,
                               int32_t bias, cudaStream_t s)
{
    const uint32_t wp = (w + (uint32_t)k / 2 + 15u) & ~15u;
    const size_t nb_pad = (size_t)wp * h * 3;
    uint8_t *tin = nullptr, *tout = nullptr;
    cudaError_t e = cudaMallocAsync(&tin, nb_pad, s);
    if (e != cudaSuccess) return e;
    e = cudaMallocAsync(&tout, nb_pad, s);
    if (e != cudaSuccess) {
        cudaFreeAsync(tin, s);
        return e;
    }
    // (one 96-thread CTA per row)
    if (e == cudaSuccess) e = geom_repitch(src, tin, w, h, w * 3u, wp * 3u, s);  // (the copy engine's 2-D path is several times slower)
    if (e == cudaSuccess) {
        launch(conv_pad_edge_kernel, dim3(h), dim3(96), 0, s, tin, w, wp);  // (wp - w <= k/2 + 15 <= 22 pixels = 66 bytes)
        e = PPMX_LAUNCHED();
    }
    if (e == cudaSuccess) e = conv(tin, tout, wp, h, k, coef, div, bias, Band(), s);
    if (e == cudaSuccess) e = geom_repitch(tout, dst, w, h, wp * 3u, w * 3u, s);
    cudaFreeAsync(tin, s);
    cudaFreeAsync(tout, s);
    return e;
}

cudaError_t conv(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div,
                 int32_t bias, const Band &band, cudaStream_t s)
{
    if (k < 1 || k > CONV_MAXK || !(k & 1) || div < 1) return cudaErrorInvalidValue;
    if (!w || !h) return cudaSuccess;
    if (band.full_h) {  // a row band: the k/2 rows beyond either cut must be reachable through the halo pointers
        const uint32_t r = (uint32_t)k / 2;
        if ((uint64_t)band.y0 + h > band.full_h) return cudaErrorInvalidValue;
        if (band.y0 > 0 && r > 0 && (!band.top || band.halo < r)) return cudaErrorInvalidValue;
        if (band.y0 + h < band.full_h && r > 0 && (!band.bottom || band.halo < r)) return cudaErrorInvalidValue;
    }
    const RowSource rs = make_row_source(src, h, band);

    bool s8 = true;
    int64_t sum_abs = 0;
    for (int i = 0; i < k * k; i++) {
        if (coef[i] < -128 || coef[i] > 127) s8 = false;
        sum_abs += coef[i] < 0 ? -(int64_t)coef[i] : coef[i];
    }
    ConvRound rnd;
    const bool fast_layout = (w % 16u) == 0 && aligned16(src) && (!band.top || aligned16(band.top)) &&
                             (!band.bottom || aligned16(band.bottom)) && PPMX_VARIANT != 1;
    uint32_t bm, bc;
    if (fast_layout && k >= 5 && k <= 11 && PPMX_VARIANT != 7 && aligned16(dst) && box_constants(coef, k, div, bias, &bm, &bc)) {
        // running sums: the cost does not depend on k (3R <= 15 halo columns come from one neighbouring lane)
        if (k == 5) return conv_box<5>(rs, dst, w, h, bm, bc, s);
        if (k == 7) return conv_box<7>(rs, dst, w, h, bm, bc, s);
        if (k == 9) return conv_box<9>(rs, dst, w, h, bm, bc, s);
        return conv_box<11>(rs, dst, w, h, bm, bc, s);
    }
    const bool rnd_ok = make_conv_round(sum_abs, div, bias, &rnd);
    [[maybe_unused]] const int strip_rh = PPMX_VARIANT >= 15 && PPMX_VARIANT <= 19 ? PPMX_VARIANT - 14 : 0;  // strip geometry (tuning build)
    if (rnd_ok && fast_layout && (k == 5 || k == 7) && PPMX_VARIANT != 7 && PPMX_VARIANT != 2 && PPMX_VARIANT != 14 && aligned16(dst)) {
        // rank-1 whose column sums fit 16 bits (ppmx_conv_sep.cu): packed pairs, dp2a along the row
        cudaError_t e = cudaSuccess;
        if (conv_sep16(rs, dst, w, h, k, coef, div, bias, rnd, strip_rh, s, &e)) return e;
    }
    if (rnd_ok && fast_layout && (k == 5 || k == 7) && PPMX_VARIANT != 7 && PPMX_VARIANT != 2 && aligned16(dst)) {
        // rank-1 (only the two factors must be small, not their products): sliding vertical words, 32-bit column sums
        cudaError_t e = cudaSuccess;
        if (k == 5 ? conv_sep<5>(rs, dst, w, h, coef, rnd, s, &e) : conv_sep<7>(rs, dst, w, h, coef, rnd, s, &e)) return e;
    }
    if (k == 3 && !s8 && rnd_ok && fast_layout && aligned16(dst) && PPMX_VARIANT != 7) {
        // coefficients beyond a signed byte (-16320 .. 16319 = 128 * (-127 .. 127) + (-64 .. 63)): the strip kernel splits them into two dp4a chains
        bool splittable = true;
        for (int i = 0; i < 9; i++) splittable = splittable && coef[i] >= -16320 && coef[i] <= 16319;
        if (splittable) return conv3_strip(rs, dst, w, h, coef, rnd, div, bias, s);
    }
    if ((k == 5 || k == 7) && fast_layout && aligned16(dst) && rnd_ok && PPMX_VARIANT != 7 && PPMX_VARIANT != 14) {  // (coefficients -16320 .. 16319)
        // dense 5x5 / 7x7: two dp4a per tap column on the vertical words (ppmx_conv_sep.cu)
        cudaError_t e = cudaSuccess;
        if (conv_dense_strip(rs, dst, w, h, k, coef, rnd, strip_rh, s, &e)) return e;
    }
    if (s8 && (k == 3 || k == 5 || k == 7) && fast_layout && aligned4(dst) && rnd_ok) {
        // variant 7 keeps the row-wise (planar dp4a) kernel for 3x3; the strip kernel is the default
        if (k == 3 && PPMX_VARIANT != 7 && aligned16(dst)) return conv3_strip(rs, dst, w, h, coef, rnd, div, bias, s);
        if (k == 3) return conv_fast<3>(rs, dst, w, h, coef, rnd, s);
        if (k == 5) return conv_fast<5>(rs, dst, w, h, coef, rnd, s);
        return conv_fast<7>(rs, dst, w, h, coef, rnd, s);
    }
    if (k >= 9 && fast_layout && aligned16(dst) && rnd_ok && PPMX_VARIANT != 7 && PPMX_VARIANT != 14 && PPMX_VARIANT != 2) {
        // rank-1 9x9 .. 15x15 blurs: column sums in shared memory, dp2a along the row (ppmx_conv_vw.cu)
        cudaError_t e = cudaSuccess;
        if (conv_vwsep(rs, dst, w, h, k, coef, div, bias, rnd, s, &e)) return e;
    }
    if (s8 && k >= 9 && fast_layout && aligned16(dst) && rnd_ok && PPMX_VARIANT != 7 && PPMX_VARIANT != 14) {
        // dense 9x9 .. 15x15: vertical words in shared memory (ppmx_conv_vw.cu); it was the scalar kernel
        cudaError_t e = cudaSuccess;
        if (conv_vw(rs, dst, w, h, k, coef, rnd, s, &e)) return e;
    }
    if (k == 3 && rnd_ok && !fast_layout && w >= 16 && (size_t)w * h >= 4096 && (PPMX_VARIANT == 0 || (PPMX_VARIANT >= 20 && PPMX_VARIANT <= 22))) {
        // 3x3 at any width / alignment (whole rasters and row bands alike): the strip kernel's unaligned form
        bool ok16 = true;
        for (int i = 0; i < 9; i++) ok16 = ok16 && coef[i] >= -16320 && coef[i] <= 16319;
        if (ok16) return conv3_strip(rs, dst, w, h, coef, rnd, div, bias, s, true);
    }
    // a whole raster that only lacks the layout for the vector kernels goes through a padded copy (variant 1 = never)
    const bool layout_only = ((w % 16u) != 0 || !aligned16(src) || !aligned16(dst)) && !band.full_h && PPMX_VARIANT != 1 &&
                             PPMX_VARIANT != 7 && w >= 16 && (size_t)w * h >= 4096 &&
                             (k <= 7 || s8 || k >= 9);  // a vector kernel may exist (the padded call falls through to the scalar kernel if not)
    if (layout_only) return conv_padded(src, dst, w, h, k, coef, div, bias, s);
    ConvCoefGeneric cf;
    for (int i = 0; i < CONV_MAXK * CONV_MAXK; i++) cf.c[i] = i < k * k ? coef[i] : 0;
    dim3 grid((w + CONV_TW - 1) / CONV_TW, (h + CONV_TH - 1) / CONV_TH);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
    size_t smem = (size_t)(CONV_TH + k - 1) * (CONV_TW + k - 1) * 3;
    launch(conv_kernel, dim3(grid), dim3(256), smem, s, rs, dst, w, k, div, bias, cf);
    return PPMX_LAUNCHED();
}


}  // namespace ppmx
