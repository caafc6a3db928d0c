// ppmx_conv_ua.cu -- EXTENSION (no reference counterpart, parity unpinned): the 3x3 strip kernel for rasters of ANY width at ANY
// pointer alignment (a width that is no multiple of 16 pixels, odd pointers: real photo sizes, raster b of a batch of such).
// Part of libppmx_gpu.so; conventions in ppmx_common.cuh, rounding in ppmx_conv.cuh, the arithmetic is conv3_strip_body's
// (ppmx_conv.cu): vertical dp4a on the interleaved raster, two output rows per set of vertical words.
//
// What is different is how bytes get in and out.  A warp owns 32 consecutive 16-byte chunks of a row (chunks are counted from
// the ROW's first byte, so the mirror rule and the 3-byte pixel stride stay where they are) over a strip of 4 output rows:
//   in:  for each of the 6 source rows the warp copies the 520-byte window [16 c0 - 4, 16 c0 + 516) of that row into shared
//        memory with cp.async of 4 bytes, lane l taking the words l, l+32, ...: every copy instruction is one fully coalesced
//        128-byte request, and the WORD part of the row's misalignment disappears in the addresses.  What remains is a byte
//        shift of 0..3, the same for the whole warp and row: a thread reads its window with two conflict-free 16-byte loads and
//        applies six funnel shifts.  (The kernel this replaces loaded three aligned vectors per thread and row and selected
//        words with 15 conditional moves; its first and last chunk of a row went byte by byte, stalling their warps.)
//        Words that lie wholly outside the row are never read (first / last row of an allocation); the three mirrored bytes at a
//        row's ends are patched into the staged window by three lanes.
//   out: the 16 result bytes of the 32 threads are one run of the destination row; they are laid down in shared memory and leave
//        as coalesced 4-byte stores (lane l writes the words l, l+32, ... of the run, each assembled by one funnel shift), the
//        up to 3 bytes at either end of the run one by one.
#include "ppmx_conv.cuh"

namespace ppmx {

constexpr int UA_PITCH = 656;   // staged source row: 520-byte window + 3 bytes of shift = 131 words; inner warps copy 160 words (no predicate)
constexpr int UA_OPITCH = 544;  // staged result row: 512 bytes + the word behind it

__device__ __forceinline__ void cp_async4(uint32_t smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// the same at a constant byte offset OFF of both addresses (an immediate in the instruction: no address arithmetic per copy)
template <int OFF>
__device__ __forceinline__ void cp_async4_at(uint32_t smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %2], 4;" ::"r"(smem_dst), "l"(gmem_src), "n"(OFF) : "memory");
}
// (Measured and dropped: shifting the results in registers before they are staged -- one shuffle + four funnel shifts per lane,
// then ONE shared-memory load per word instead of two -- gave 0.541 against 0.549 at 4090^2: 7 more registers, one CTA fewer per SM.)
// `nb` bytes from shared memory (16-byte aligned, 4 bytes of slack behind the run) to any global address, by one warp: whole
// 4-byte words of the destination as coalesced stores (lane l: words l, l+32, ...), the up to 3 bytes in front of the first
// and behind the last word by lanes 0..2 and 3..5
__device__ __forceinline__ void ua_store_run(uint8_t *g, const uint8_t *srow, uint32_t nb, uint32_t lane)
{
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 3u);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(srow) + lane;
    uint32_t hb, nw, tb;
    if (nb == 512u) {  // a full run (all warps but a row's last): only the last word of the last lane may be missing
        hb = (4u - a) & 3u;
        nw = hb ? 127u : 128u;
        tb = hb ? 4u - hb : 0u;
        uint32_t *gw = reinterpret_cast<uint32_t *>(g + hb) + lane;
        const uint32_t sh = 8u * hb;
#pragma unroll
        for (uint32_t t = 0; t < 3u; t++) gw[32u * t] = __funnelshift_r(sw[32u * t], sw[32u * t + 1], sh);
        if (lane < 31u || !hb) gw[96] = __funnelshift_r(sw[96], sw[97], sh);
    } else {
        hb = min(nb, (4u - a) & 3u);
        nw = (nb - hb) >> 2;
        tb = nb - hb - 4u * nw;
        uint32_t *gw = reinterpret_cast<uint32_t *>(g + hb) + lane;
        const uint32_t sh = 8u * hb;
#pragma unroll
        for (uint32_t t = 0; t < 4u; t++)
            if (lane + 32u * t < nw) gw[32u * t] = __funnelshift_r(sw[32u * t], sw[32u * t + 1], sh);
    }
    const bool tail = lane >= 3u;
    const uint32_t k = tail ? lane - 3u : lane, pos = tail ? hb + 4u * nw + k : k;
    if (k < (tail ? tb : hb) && lane < 6u) g[pos] = srow[pos];
}

template <int MODE, bool WIDE, int RH>
__global__ void __launch_bounds__(128) conv3_ua_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t nchunks, uint32_t row_bytes,
                                                       const Conv3Coef cf, const ConvRound rnd)
{
    pdl_trigger();
    constexpr int NR = RH + 2, NG = RH / 2;  // source rows of a strip; row pairs = cp.async groups
    __shared__ __align__(16) uint8_t s_in[4][NR][UA_PITCH];
    __shared__ __align__(16) uint8_t s_out[4][2][UA_OPITCH];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t c0 = blockIdx.x * 128u + warp * 32u;  // the warp's first chunk
    if (c0 >= nchunks) return;                            // (a whole warp)
    const int ys = blockIdx.y * RH;                       // first output row of the strip, band-local
    const size_t pitch = row_bytes;
    const bool inner = ys >= 1 && ys + RH + 1 <= rs.h;  // source rows ys-1 .. ys+RH all in the own band
    const int gy0 = rs.y0 + ys;
    // the window reaches beyond the row's first / last byte; `far`: 640 bytes from the window's start still lie inside the row
    const bool left = c0 == 0, right = 16u * c0 + 516u > row_bytes, far = 16u * c0 + 640u <= row_bytes;
    const int64_t base_b = (int64_t)16 * c0 - 4;                     // row byte of the window's first byte
    uint8_t(*sin)[UA_PITCH] = s_in[warp];
    pdl_wait();

    // ---- in: the row windows -> shared memory, 4 bytes per copy, coalesced; rows 0..3 are group 0, every further pair one more ----
    uint32_t shift[NR];
    const uint8_t *rowp = rs.own + (ptrdiff_t)(ys - 1) * (ptrdiff_t)pitch;  // (inner strips: plain pointer steps)
    const uint32_t sdst0 = (uint32_t)__cvta_generic_to_shared(sin[0]) + 4u * lane;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        if (!inner) rowp = rs.row(gy0 - 1 + r, pitch);
        const uintptr_t g = reinterpret_cast<uintptr_t>(rowp) + (uintptr_t)base_b;
        shift[r] = 8u * (uint32_t)(g & 3u);
        const uint8_t *gsrc = reinterpret_cast<const uint8_t *>(g & ~(uintptr_t)3) + 4u * lane;  // this lane's first word of the window
        const uint32_t sdst = sdst0 + (uint32_t)(r * UA_PITCH);
        if (!left && far) {  // 5 x 32 words, the last 29 of them unused
            cp_async4_at<0>(sdst, gsrc);
            cp_async4_at<128>(sdst, gsrc);
            cp_async4_at<256>(sdst, gsrc);
            cp_async4_at<384>(sdst, gsrc);
            cp_async4_at<512>(sdst, gsrc);
        } else {  // only words holding at least one byte of the row
            const uintptr_t lo = reinterpret_cast<uintptr_t>(rowp) & ~(uintptr_t)3;
            const uintptr_t hi = (reinterpret_cast<uintptr_t>(rowp) + row_bytes - 1u) & ~(uintptr_t)3;
#pragma unroll
            for (uint32_t t = 0; t < 5u; t++) {
                const uintptr_t wa = reinterpret_cast<uintptr_t>(gsrc) + 128u * t;
                if (lane + 32u * t < 131u && wa >= lo && wa <= hi) cp_async4(sdst + 128u * t, reinterpret_cast<const void *>(wa));
            }
        }
        if (inner) rowp += pitch;
        if (r >= 3 && (r & 1)) asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (left || right) {  // pixel -1 mirrors to pixel 0, pixel W to pixel W-1: three bytes each, per row
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (lane < 3u * NR) {
            const uint32_t r = lane / 3u, k = lane - 3u * r;
            uint32_t sr = shift[0] >> 3;  // (shift[r]: selected without dynamic indexing)
#pragma unroll
            for (int q = 1; q < NR; q++) sr = r == (uint32_t)q ? (shift[q] >> 3) : sr;
            uint8_t *srow = sin[r];
            if (left) srow[sr + 1u + k] = srow[sr + 4u + k];
            if (right) {
                const uint32_t e = (uint32_t)((int64_t)row_bytes - base_b) + sr;  // staged position of the row's end
                srow[e + k] = srow[e - 3u + k];
            }
        }
        __syncwarp();
    }

    // ---- the strip's arithmetic (conv3_strip_body): vertical words of row pairs, three dp4a per output byte ----
    auto window = [&](int r, uint32_t(&w6)[6]) {
        const uint4 *p = reinterpret_cast<const uint4 *>(sin[r] + 16u * lane);
        const uint4 x = p[0], y = p[1];
        const uint32_t sh = shift[r];
        w6[0] = __funnelshift_r(x.x, x.y, sh);
        w6[1] = __funnelshift_r(x.y, x.z, sh);
        w6[2] = __funnelshift_r(x.z, x.w, sh);
        w6[3] = __funnelshift_r(x.w, y.x, sh);
        w6[4] = __funnelshift_r(y.x, y.y, sh);
        w6[5] = __funnelshift_r(y.y, y.z, sh);
    };
    asm volatile("cp.async.wait_group %0;" ::"n"(NG - 1) : "memory");  // rows 0..3 (group 0) before the first pair
    __syncwarp();
    uint32_t tlo[6], thi[6];
    {
        uint32_t a[6], b[6];
        window(0, a);
        window(1, b);
#pragma unroll
        for (int wc = 0; wc < 6; wc++) {
            tlo[wc] = __byte_perm(a[wc], b[wc], 0x5140);
            thi[wc] = __byte_perm(a[wc], b[wc], 0x7362);
        }
    }
    const uint32_t run = min(512u, row_bytes - 16u * c0);
    uint8_t(*sout)[UA_OPITCH] = s_out[warp];
#pragma unroll
    for (int g = 0; g < NG; g++) {
        if (!inner && ys + 2 * g >= rs.h) return;
        if (g == 1) asm volatile("cp.async.wait_group %0;" ::"n"(NG >= 2 ? NG - 2 : 0) : "memory");
        if (g == 2) asm volatile("cp.async.wait_group %0;" ::"n"(NG >= 3 ? NG - 3 : 0) : "memory");
        if (g == 3) asm volatile("cp.async.wait_group %0;" ::"n"(NG >= 4 ? NG - 4 : 0) : "memory");
        if (g >= 1) __syncwarp();
        uint32_t ulo[6], uhi[6];
        {
            uint32_t a[6], b[6];
            window(2 * g + 2, a);
            window(2 * g + 3, b);
#pragma unroll
            for (int wc = 0; wc < 6; wc++) {
                ulo[wc] = __byte_perm(a[wc], b[wc], 0x5140);
                uhi[wc] = __byte_perm(a[wc], b[wc], 0x7362);
            }
        }
        uint32_t V[24];  // V[i] = byte column 16 cx - 4 + i: rows y-1, y, y+1, y+2 in its four bytes
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
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int32_t accA[4], accB[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = 4 * b + j + 4;
                accA[j] = dp4a_u8s8(V[c + 3], cf.a[2], dp4a_u8s8(V[c], cf.a[1], dp4a_u8s8(V[c - 3], cf.a[0], rnd.start)));
                accB[j] = dp4a_u8s8(V[c + 3], cf.b[2], dp4a_u8s8(V[c], cf.b[1], dp4a_u8s8(V[c - 3], cf.b[0], rnd.start)));
                if (WIDE) {  // the high parts of the coefficients: a second dp4a chain, weighted 128
                    accA[j] += dp4a_u8s8(V[c + 3], cf.ah[2], dp4a_u8s8(V[c], cf.ah[1], dp4a_u8s8(V[c - 3], cf.ah[0], 0))) << 7;
                    accB[j] += dp4a_u8s8(V[c + 3], cf.bh[2], dp4a_u8s8(V[c], cf.bh[1], dp4a_u8s8(V[c - 3], cf.bh[0], 0))) << 7;
                }
            }
            oa[b] = rnd.template pack4<MODE>(accA[0], accA[1], accA[2], accA[3]);
            ob[b] = rnd.template pack4<MODE>(accB[0], accB[1], accB[2], accB[3]);
        }
        // ---- out: the warp's run of the two destination rows ----
        reinterpret_cast<uint4 *>(sout[0])[lane] = make_uint4(oa[0], oa[1], oa[2], oa[3]);
        reinterpret_cast<uint4 *>(sout[1])[lane] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        __syncwarp();
        uint8_t *o0 = dst + (size_t)(ys + 2 * g) * pitch + (size_t)c0 * 16;
        ua_store_run(o0, sout[0], run, lane);
        if (inner || ys + 2 * g + 1 < rs.h) ua_store_run(o0 + pitch, sout[1], run, lane);
        __syncwarp();
    }
}

template <int RH>
static cudaError_t conv3_ua_rh(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const Conv3Coef &cf, const ConvRound &rnd,
                               int mode, bool wide, cudaStream_t s)
{
    const uint32_t row_bytes = w * 3u, nch = (row_bytes + 15u) / 16u;
    dim3 grid((nch + 127) / 128, (h + RH - 1) / RH);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
#define PPMX_CONV3_UA(MODE)                                                                                           \
    do {                                                                                                              \
        if (wide) launch(conv3_ua_kernel<MODE, true, RH>, grid, dim3(128), 0, s, rs, dst, nch, row_bytes, cf, rnd);   \
        else launch(conv3_ua_kernel<MODE, false, RH>, grid, dim3(128), 0, s, rs, dst, nch, row_bytes, cf, rnd);       \
    } while (0)
    if (mode == 0) PPMX_CONV3_UA(0);
    else if (mode == 1) PPMX_CONV3_UA(1);
    else if (mode == 3) PPMX_CONV3_UA(3);
    else PPMX_CONV3_UA(2);
#undef PPMX_CONV3_UA
    return PPMX_LAUNCHED();
}

cudaError_t conv3_ua_launch(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const Conv3Coef &cf, const ConvRound &rnd,
                            int mode, bool wide, cudaStream_t s)
{
#ifdef PPMX_TUNING
    if (PPMX_VARIANT == 21) return conv3_ua_rh<8>(rs, dst, w, h, cf, rnd, mode, wide, s);
    if (PPMX_VARIANT == 22) return conv3_ua_rh<2>(rs, dst, w, h, cf, rnd, mode, wide, s);
#endif
    return conv3_ua_rh<4>(rs, dst, w, h, cf, rnd, mode, wide, s);
}

}  // namespace ppmx
