// ppmx_geometry.cu -- flip (ref:898-911) and the exact rotations by 90 / 180 / 270 degrees (ref:714-725): pure byte moves.
// Part of libppmx_gpu.so; see ppmx_common.cuh for conventions ("ref:N" = /root/reference/ppmx-edward.c line N).
#include "ppmx_common.cuh"

namespace ppmx {

// ------------------------------------------------------------------------------------------
// flip  (ref:898-911).  The reference swaps in place; here dst != src, same bytes out.
// ------------------------------------------------------------------------------------------

// vertical: row y of dst = row h-1-y of src; T = widest type the row pitch and pointers allow
template <typename T>
__global__ void __launch_bounds__(256) flipv_kernel(const T *__restrict__ src, T *__restrict__ dst,
                                                    uint32_t row_elems, uint32_t h, size_t n)
{
    pdl_trigger();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t y, e;
    if (n <= 0xFFFFFFFFull) {  // 32-bit division is several times cheaper than 64-bit
        y = (uint32_t)i / row_elems;
        e = (uint32_t)i - y * row_elems;
    } else {
        y = (uint32_t)(i / row_elems);
        e = (uint32_t)(i - (size_t)y * row_elems);
    }
    pdl_wait();  // everything above is index arithmetic; global memory is touched only below
    dst[i] = src[(size_t)(h - 1 - y) * row_elems + e];
}

// the same for 16-byte rows, four rows per thread (same vector of rows y .. y+3: no division, four loads in flight before the
// first store): 0.95 -> of the HBM roofline for the one-vector form at 4096^2
__global__ void __launch_bounds__(256) flipv_rows4_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint32_t row_elems,
                                                          uint32_t h)
{
    pdl_trigger();
    const uint32_t e = blockIdx.x * 256u + threadIdx.x, y0 = blockIdx.y * 4u;
    if (e >= row_elems) return;
    const uint32_t n = min(4u, h - y0);
    const uint4 *s = src + (size_t)(h - 1u - y0) * row_elems + e;  // source row of output row y0; the next ones lie BEFORE it
    uint4 *d = dst + (size_t)y0 * row_elems + e;
    pdl_wait();
    uint4 v[4];
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++)
        if (k < n) v[k] = __ldg(s - (size_t)k * row_elems);
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++)
        if (k < n) d[(size_t)k * row_elems] = v[k];
}

// horizontal, any pixel size / alignment: one byte per thread
__global__ void __launch_bounds__(256) fliph_generic_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                            uint32_t w, uint32_t h, int bpp)
{
    PDL_PROLOGUE();
    const size_t row = (size_t)w * bpp, n = row * h, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        size_t y = i / row;
        uint32_t b = (uint32_t)(i - y * row), x = b / bpp, c = b - x * bpp;
        dst[i] = src[y * row + (size_t)(w - 1 - x) * bpp + c];
    }
}

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[12], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

// reverse the order of 16 RGB pixels held in 12 words
__device__ __forceinline__ void reverse16px(const uint32_t (&in)[12], uint32_t (&out)[12])
{
#pragma unroll
    for (int k = 0; k < 12; k++) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int o = 4 * k + j;                       // output byte
            int i = 3 * (15 - o / 3) + (o % 3);      // input byte: same channel of the mirrored pixel
            v |= byte_of(in, i) << (8 * j);
        }
        out[k] = v;
    }
}

// horizontal RGB8, w % 16 == 0, aligned: a thread moves one 16-pixel group (48 B) to its mirror slot
__global__ void __launch_bounds__(256) fliph_rgb16_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                          uint32_t groups_per_row, size_t ngroups)
{
    pdl_trigger();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= ngroups) return;
    size_t y;
    uint32_t g;
    if (ngroups <= 0xFFFFFFFFull) {
        y = (uint32_t)i / groups_per_row;
        g = (uint32_t)i - (uint32_t)y * groups_per_row;
    } else {
        y = i / groups_per_row;
        g = (uint32_t)(i - y * groups_per_row);
    }
    pdl_wait();  // everything above is index arithmetic; global memory is touched only below
    const uint4 *p = src + 3 * (y * groups_per_row + (groups_per_row - 1 - g));
    const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    const uint32_t in[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    uint32_t o[12];
    reverse16px(in, o);
    uint4 *q = dst + 3 * i;
    q[0] = make_uint4(o[0], o[1], o[2], o[3]);
    q[1] = make_uint4(o[4], o[5], o[6], o[7]);
    q[2] = make_uint4(o[8], o[9], o[10], o[11]);
}

// Rows of any length at any alignment (widths that are no multiple of 16, odd pointers): one grid row per raster
// row, one thread per ALIGNED destination word.  Vertical: the four source bytes come out of two aligned source
// words by a funnel shift (source and destination rows are misaligned differently); horizontal: four byte loads
// from the mirrored pixels.  Words cut by the row's ends are written byte by byte.  No 64-bit division anywhere.
template <bool HFLIP, int BPP, bool VFLIP = !HFLIP>
__global__ void __launch_bounds__(256) flip_rows_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t w,
                                                        uint32_t h)
{
    PDL_PROLOGUE();
    const uint32_t y = blockIdx.y, row = w * BPP;
    uint8_t *D0 = dst + (size_t)y * row;
    const uint8_t *S0 = src + (size_t)(VFLIP ? h - 1u - y : y) * row;  // (both flips together = 180 degrees, ref:721)
    const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(D0) & 3u), nwords = (a0 + row + 3u) / 4u;
    for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < nwords; k += gridDim.x * 256u) {
        const int b0 = (int)(4u * k) - (int)a0;  // first row byte of this destination word
        const bool full = b0 >= 0 && (uint32_t)b0 + 4u <= row;
        if (!HFLIP && full) {
            const uint8_t *q = S0 + b0;
            const uint32_t m = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3u);
            const uint32_t *base = reinterpret_cast<const uint32_t *>(q - m);
            const uint32_t w0 = __ldg(base), w1 = m ? __ldg(base + 1) : 0u;
            *reinterpret_cast<uint32_t *>(D0 + b0) = __funnelshift_r(w0, w1, 8u * m);
            continue;
        }
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int b = b0 + j;
            if (b < 0 || (uint32_t)b >= row) continue;
            uint32_t sb = (uint32_t)b;
            if (HFLIP) {
                const uint32_t x = (uint32_t)b / BPP, c = (uint32_t)b - x * BPP;
                sb = (w - 1u - x) * BPP + c;
            }
            const uint32_t byte = S0[sb];
            if (full) v |= byte << (8 * j);
            else D0[b] = (uint8_t)byte;
        }
        if (full) *reinterpret_cast<uint32_t *>(D0 + b0) = v;
    }
}

cudaError_t flip(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int bpp, int vertical, cudaStream_t s)
{
    if (!w || !h) return cudaSuccess;
    size_t row = (size_t)w * bpp;
    if (vertical) {
        if (row % 16 == 0 && aligned16(src) && aligned16(dst)) {
            size_t n = row / 16 * h;
            if (h <= 4u * 65535u && PPMX_VARIANT != 2)
                launch(flipv_rows4_kernel, dim3((unsigned)((row / 16 + 255) / 256), (h + 3u) / 4u), dim3(256), 0, s,
                       reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), (uint32_t)(row / 16), h);
            else
                launch(flipv_kernel<uint4>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s,
                       reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), (uint32_t)(row / 16), h, n);
        } else if (row % 4 == 0 && aligned4(src) && aligned4(dst)) {
            size_t n = row / 4 * h;
            launch(flipv_kernel<uint32_t>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s,
                   reinterpret_cast<const uint32_t *>(src), reinterpret_cast<uint32_t *>(dst), (uint32_t)(row / 4), h, n);
        } else if (bpp == 3 && PPMX_VARIANT == 0 && (size_t)w * h >= 4096 && h <= 65535u * 64u) {
            // rows at any alignment (width no multiple of 16): the tile kernel of ppmx_fused.cu
            GeomOp go = {};
            go.rev_y = 1;
            return geom_point(src, dst, w, h, 0, go, s);
        } else if (h <= 65535u && PPMX_VARIANT != 1) {
            const unsigned gx = (unsigned)((row / 4 + 1 + 255) / 256);
            if (bpp == 3) launch(flip_rows_kernel<false, 3>, dim3(gx < 64u ? gx : 64u, h), dim3(256), 0, s, src, dst, w, h);
            else launch(flip_rows_kernel<false, 1>, dim3(gx < 64u ? gx : 64u, h), dim3(256), 0, s, src, dst, w, h);
        } else {
            launch(flipv_kernel<uint8_t>, dim3((unsigned)((row * h + 255) / 256)), dim3(256), 0, s, src, dst, (uint32_t)row, h,
                   row * h);
        }
    } else {
        if (bpp == 3 && (w % 16u) == 0 && aligned16(src) && aligned16(dst)) {
            size_t n = (size_t)(w / 16u) * h;
            launch(fliph_rgb16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s,
                   reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), w / 16u, n);
        } else if (bpp == 3 && PPMX_VARIANT == 0 && (size_t)w * h >= 4096 && h <= 65535u * 64u) {
            GeomOp go = {};
            go.rev_x = 1;
            return geom_point(src, dst, w, h, 0, go, s);
        } else if (h <= 65535u && (bpp == 3 || bpp == 1) && PPMX_VARIANT != 1) {
            const unsigned gx = (unsigned)((row / 4 + 1 + 255) / 256);
            if (bpp == 3) launch(flip_rows_kernel<true, 3>, dim3(gx < 64u ? gx : 64u, h), dim3(256), 0, s, src, dst, w, h);
            else launch(flip_rows_kernel<true, 1>, dim3(gx < 64u ? gx : 64u, h), dim3(256), 0, s, src, dst, w, h);
        } else {
            launch(fliph_generic_kernel, dim3(wave_grid(row * h, 256, 8)), dim3(256), 0, s, src, dst, w, h, bpp);
        }
    }
    return PPMX_LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// rotate 90 / 180 / 270  (ref:714-725): pure byte moves
// ------------------------------------------------------------------------------------------

// 180 degrees reverses the whole pixel sequence: out[n-1-p] = in[p]  (ref:721)
__global__ void __launch_bounds__(256) reverse_pixels_kernel(const uint8_t *__restrict__ src,
                                                             uint8_t *__restrict__ dst, size_t npix)
{
    PDL_PROLOGUE();
    const size_t n = npix * 3, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        size_t p = i / 3;
        uint32_t c = (uint32_t)(i - p * 3);
        dst[i] = src[(npix - 1 - p) * 3 + c];
    }
}

// 90 / 270: transpose through a 32 x 32 pixel shared-memory tile.
//   90:  out[x][h-1-y] = in[y][x]   (ref:717)      out is h wide, w tall
//   270: out[w-1-x][y] = in[y][x]   (ref:725)
constexpr int RT = 32;            // tile edge in pixels
[[maybe_unused]] constexpr int RT_PITCH = RT * 3 + 4;  // bytes; +4 keeps column reads off a single bank

template <bool CW>
__global__ void __launch_bounds__(256) rotate_transpose_kernel(const uint8_t *__restrict__ src,
                                                               uint8_t *__restrict__ dst, uint32_t w, uint32_t h)
{
    PDL_PROLOGUE();
    __shared__ uint8_t tile[RT][RT_PITCH];
    const uint32_t tx0 = blockIdx.x * RT, ty0 = blockIdx.y * RT;
    const uint32_t tw = min((uint32_t)RT, w - tx0), th = min((uint32_t)RT, h - ty0);
    const size_t in_pitch = (size_t)w * 3, out_pitch = (size_t)h * 3;

    for (uint32_t i = threadIdx.x; i < th * tw * 3; i += blockDim.x) {
        uint32_t r = i / (tw * 3), b = i - r * (tw * 3);
        tile[r][b] = src[(size_t)(ty0 + r) * in_pitch + (size_t)tx0 * 3 + b];
    }
    __syncthreads();
    // the output tile has tw rows of th pixels
    for (uint32_t i = threadIdx.x; i < tw * th * 3; i += blockDim.x) {
        uint32_t orow = i / (th * 3), ob = i - orow * (th * 3), opx = ob / 3, ch = ob - opx * 3;
        if (CW) {  // out row = x, out col = h-1-y: columns run against y
            uint32_t r = th - 1 - opx;
            size_t ocol0 = (size_t)(h - ty0 - th);
            dst[(size_t)(tx0 + orow) * out_pitch + (ocol0 + opx) * 3 + ch] = tile[r][orow * 3 + ch];
        } else {  // out row = w-1-x, out col = y
            uint32_t c = tw - 1 - orow;
            size_t orow_g = (size_t)(w - tx0 - tw) + orow;
            dst[orow_g * out_pitch + ((size_t)ty0 + opx) * 3 + ch] = tile[opx][c * 3 + ch];
        }
    }
}

// The same for any raster, 64 x 64 pixel tiles, warps own rows: a warp copies source rows into the tile 32 consecutive
// bytes at a time and writes destination rows the same way (every global access is one 32-byte sector per
// instruction); no division by a run-time value.  4090 x 4090: 0.125 (32 x 32 byte-indexed tile, variant 1) -> 0.26 of the roofline.
constexpr int RA = 64;
[[maybe_unused]] constexpr int RA_PITCH = RA * 3 + 4;  // 196 B = 49 words: consecutive tile rows start 17 banks apart

template <bool CW>
__global__ void __launch_bounds__(256) rotate_transpose_any_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                                   uint32_t w, uint32_t h)
{
    PDL_PROLOGUE();
    __shared__ uint8_t tile[RA][RA_PITCH];
    const uint32_t tx0 = blockIdx.x * RA, ty0 = blockIdx.y * RA;
    const uint32_t tw = min((uint32_t)RA, w - tx0), th = min((uint32_t)RA, h - ty0);
    const size_t in_pitch = (size_t)w * 3, out_pitch = (size_t)h * 3;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t r = warp; r < th; r += 8) {
        const uint8_t *p = src + (size_t)(ty0 + r) * in_pitch + (size_t)tx0 * 3;
        for (uint32_t b = lane; b < tw * 3u; b += 32) tile[r][b] = p[b];
    }
    __syncthreads();
    // the output tile has tw rows of th pixels
    for (uint32_t orow = warp; orow < tw; orow += 8) {
        uint8_t *q;
        if (CW) q = dst + (size_t)(tx0 + orow) * out_pitch + (size_t)(h - ty0 - th) * 3;       // out[x][h-1-y], ref:717
        else q = dst + (size_t)(w - tx0 - tw + orow) * out_pitch + (size_t)ty0 * 3;            // out[w-1-x][y], ref:725
        const uint32_t col = CW ? orow : tw - 1u - orow;  // the source column this destination row is made of
        for (uint32_t ob = lane; ob < th * 3u; ob += 32) {
            const uint32_t opx = ob / 3u, ch = ob - opx * 3u;
            q[ob] = tile[CW ? th - 1u - opx : opx][col * 3u + ch];
        }
    }
}

// Fast path (w % 16 == 0, h % 16 == 0, 16-byte aligned rasters): 64 x 64 pixel tiles.
//   phase 1: a thread loads 16 pixels of one source row (3 x 16 B), widens them to one word per
//            pixel (r g b x) and stores 4 x 16 B into a swizzled shared tile (no bank conflicts);
//   phase 2: a lane reads one source column over 16 source rows, one word per row (conflict
//            free), repacks the 16 pixels to 48 B and writes them as three 16-byte stores into
//            the destination row that column became; 4 lanes complete 192 contiguous bytes.
[[maybe_unused]] constexpr int XT = 64;

__device__ __forceinline__ uint32_t xt_slot(uint32_t row, uint32_t chunk)
{
    // 16-byte chunk index inside a 256-byte tile row, swizzled so that both phases spread over all
    // banks: phase 1 stores rows (r, r+1) x 4 quarter rows at once, phase 2 reads rows 16 apart
    return row * 64u + ((chunk ^ (((chunk >> 3) & 1u) << 1) ^ (row & 1u) ^ (((row >> 4) & 3u) << 1)) << 2);
}

// XT_NT horizontally adjacent tiles per CTA (vertical pairs, i.e. longer contiguous WRITES, measured
// 8 % slower: long contiguous reads matter more): the loads of ALL of them are issued up front, so
// the second tile's DRAM latency hides behind the first tile's shared-memory phase and stores.
// BAND > 0: blockIdx.x = tile_x * BAND + (tile row inside a band of BAND tile rows), blockIdx.y = band; CTAs
// then walk BAND tiles down before stepping right, which lengthens the contiguous run written per
// destination row while it is "hot" (shift arithmetic only).  BAND = 0: plain 2-D grid.
template <bool CW, int XT_NT, int MINB, int BAND>
__global__ void __launch_bounds__(256, MINB) rotate_transpose64_kernel(const uint8_t *__restrict__ src,
                                                                 uint8_t *__restrict__ dst, uint32_t w, uint32_t h)
{
    pdl_trigger();
    __shared__ __align__(16) uint32_t tile[XT * 64];
    const uint32_t bx = BAND ? blockIdx.x / BAND : blockIdx.x;
    const uint32_t by = BAND ? blockIdx.y * BAND + blockIdx.x % BAND : blockIdx.y;
    const uint32_t ty0 = by * XT;
    if (ty0 >= h) return;
    const size_t in_pitch = (size_t)w * 3, out_pitch = (size_t)h * 3;
    const uint32_t row = threadIdx.x >> 2, q = threadIdx.x & 3u;               // phase 1 role
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t col = 8u * warp + (lane >> 2), j = lane & 3u;              // phase 2 role
    pdl_wait();  // everything above is index arithmetic; global memory is touched only below

    uint4 ld[XT_NT][3];
    bool have[XT_NT];
#pragma unroll
    for (int t = 0; t < XT_NT; t++) {
        const uint32_t y = ty0 + row, x0 = (bx * XT_NT + t) * XT + 16u * q;
        have[t] = (y < h && x0 < w);
        if (have[t]) {
            const uint4 *p = reinterpret_cast<const uint4 *>(src + (size_t)y * in_pitch + (size_t)x0 * 3);
            ld[t][0] = __ldg(p);
            ld[t][1] = __ldg(p + 1);
            ld[t][2] = __ldg(p + 2);
        }
    }
#pragma unroll
    for (int t = 0; t < XT_NT; t++) {
        const uint32_t tx0 = (bx * XT_NT + t) * XT;
        if (tx0 >= w) break;
        if (t > 0) __syncthreads();  // the previous tile has been read out of shared memory
        if (have[t]) {
            const uint32_t wd[12] = {ld[t][0].x, ld[t][0].y, ld[t][0].z, ld[t][0].w, ld[t][1].x, ld[t][1].y,
                                     ld[t][1].z, ld[t][1].w, ld[t][2].x, ld[t][2].y, ld[t][2].z, ld[t][2].w};
#pragma unroll
            for (int i = 0; i < 4; i++) {  // 4 pixels = 3 words -> 4 words (byte 3 of each is don't-care)
                uint4 o;
                o.x = wd[3 * i];
                o.y = __byte_perm(wd[3 * i], wd[3 * i + 1], 0x0543);
                o.z = __byte_perm(wd[3 * i + 1], wd[3 * i + 2], 0x0432);
                o.w = wd[3 * i + 2] >> 8;
                *reinterpret_cast<uint4 *>(&tile[xt_slot(row, 4u * q + i)]) = o;
            }
        }
        __syncthreads();
        // four neighbouring lanes take the four 16-row units of one source column, so together they
        // write one contiguous 192-byte piece of a destination row
        const uint32_t x = tx0 + col, y0 = ty0 + 16u * j;
        if (x < w && y0 < h) {
            uint32_t px[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t r = 16u * j + k;
                px[k] = tile[xt_slot(r, col >> 2) + (col & 3u)];
            }
            uint32_t o[12];
#pragma unroll
            for (int i = 0; i < 4; i++) {  // 4 pixels -> 3 words; CW walks the source rows backwards
                const uint32_t p0 = CW ? px[15 - 4 * i] : px[4 * i], p1 = CW ? px[14 - 4 * i] : px[4 * i + 1];
                const uint32_t p2 = CW ? px[13 - 4 * i] : px[4 * i + 2], p3 = CW ? px[12 - 4 * i] : px[4 * i + 3];
                o[3 * i] = __byte_perm(p0, p1, 0x4210);
                o[3 * i + 1] = __byte_perm(p1, p2, 0x5421);
                o[3 * i + 2] = __byte_perm(p2, p3, 0x6542);
            }
            size_t off;
            if (CW) off = (size_t)x * out_pitch + (size_t)(h - y0 - 16u) * 3;        // out[x][h-1-y], ref:717
            else off = (size_t)(w - 1u - x) * out_pitch + (size_t)y0 * 3;             // out[w-1-x][y], ref:725
            uint4 *qo = reinterpret_cast<uint4 *>(dst + off);
            qo[0] = make_uint4(o[0], o[1], o[2], o[3]);
            qo[1] = make_uint4(o[4], o[5], o[6], o[7]);
            qo[2] = make_uint4(o[8], o[9], o[10], o[11]);
        }
    }
}

// ---- 90 / 270 degrees through the bulk-copy engine (w % 16 == 0, h % 16 == 0) ------------------
// One 64 x 64 pixel tile per 128-thread CTA.  The 64 source row pieces (192 B each) come in by cp.async.bulk
// (global -> shared, completion on an mbarrier) as RAW interleaved bytes, the transposed tile leaves by
// cp.async.bulk (shared -> global) as 64 destination row pieces of 192 B: no thread issues a global load or store,
// so the L1 tag stage no longer sees 8..16 partly used lines per instruction (what held the register-path kernel
// at 0.71), and 8 CTAs per SM keep ~100 KB in flight.  In between, a lane takes ONE source column over 16 rows:
// the pixel at byte 3x of a row is cut out of two aligned words by a funnel shift whose amount is constant per
// lane; a warp's 32 lanes read 24 consecutive words of one row (conflict-free), and write three 16-byte vectors
// each into rows 208 B apart (conflict-free per quarter warp).
constexpr int XB_PITCH = 208;  // 192 B of pixels + 16 B: rows stay 16-byte aligned and spread over the banks

__device__ __forceinline__ uint32_t xb_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool CW>
__global__ void __launch_bounds__(128) rotate_bulk64_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t w,
                                                            uint32_t h)
{
    pdl_trigger();
    __shared__ __align__(128) uint8_t tin[64 * XB_PITCH];
    __shared__ __align__(128) uint8_t tout[64 * XB_PITCH];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tid = threadIdx.x, tx0 = blockIdx.x * 64u, ty0 = blockIdx.y * 64u;
    // tiles at the right / bottom edge of a raster whose sides are multiples of 16 are 16, 32 or 48 pixels wide / tall
    const uint32_t nc = min(64u, w - tx0), nr = min(64u, h - ty0);
    const size_t in_pitch = (size_t)w * 3, out_pitch = (size_t)h * 3;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xb_smem(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    if (tid == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xb_smem(&bar)), "r"(nr * nc * 3u) : "memory");
    // a bulk copy is a per-warp (uniform) instruction, so a warp issues its lanes' copies one after the other:
    // 16 per warp on all four warps instead of 32 on two halves that latency (0.71 -> 0.82 of the HBM roofline)
    const uint32_t crow = (tid >> 5) * 16u + (tid & 15u);  // the row piece this thread copies (lanes 0..15 of each warp)
    if ((tid & 31u) < 16u && crow < nr) {
        const uint8_t *g = src + (size_t)(ty0 + crow) * in_pitch + (size_t)tx0 * 3;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         xb_smem(tin + crow * XB_PITCH)),
                     "l"(g), "r"(nc * 3u), "r"(xb_smem(&bar))
                     : "memory");
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(xb_smem(&bar)),
        "r"(0)
        : "memory");

    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const uint32_t x = lane + 32u * (warp & 1u);             // source column inside the tile
    const uint32_t wcol = (3u * x) >> 2, sh = ((3u * x) & 3u) * 8u;
    const uint32_t *tin32 = reinterpret_cast<const uint32_t *>(tin);
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        const uint32_t j = (warp >> 1) + 2u * pass;          // 16-row group
        if (x >= nc || 16u * j >= nr) continue;
        uint32_t px[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t *rw = tin32 + (16u * j + k) * (XB_PITCH / 4) + wcol;
            px[k] = __funnelshift_r(rw[0], rw[1], sh);       // r g b x (the word after the last pixel is the row's padding)
        }
        uint32_t o[12];
#pragma unroll
        for (int i = 0; i < 4; i++) {  // 4 pixels -> 3 words; CW walks the source rows backwards
            const uint32_t p0 = CW ? px[15 - 4 * i] : px[4 * i], p1 = CW ? px[14 - 4 * i] : px[4 * i + 1];
            const uint32_t p2 = CW ? px[13 - 4 * i] : px[4 * i + 2], p3 = CW ? px[12 - 4 * i] : px[4 * i + 3];
            o[3 * i] = __byte_perm(p0, p1, 0x4210);
            o[3 * i + 1] = __byte_perm(p1, p2, 0x5421);
            o[3 * i + 2] = __byte_perm(p2, p3, 0x6542);
        }
        // CW: out[x][h-1-y] (ref:717): destination row x, its pixels nr-16-16j .. nr-1-16j; else out[w-1-x][y] (ref:725):
        // destination row nc-1-x, pixels 16j .. 16j+15
        const uint32_t drow = CW ? x : nc - 1u - x, dbyte = CW ? (nr - 16u - 16u * j) * 3u : 48u * j;
        uint4 *q = reinterpret_cast<uint4 *>(tout + drow * XB_PITCH + dbyte);
        q[0] = make_uint4(o[0], o[1], o[2], o[3]);
        q[1] = make_uint4(o[4], o[5], o[6], o[7]);
        q[2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the bulk engine must see the tile the threads wrote
    __syncthreads();
    if ((tid & 31u) < 16u && crow < nc) {
        uint8_t *g = CW ? dst + (size_t)(tx0 + crow) * out_pitch + (size_t)(h - ty0 - nr) * 3
                        : dst + (size_t)(w - tx0 - nc + crow) * out_pitch + (size_t)ty0 * 3;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(xb_smem(tout + crow * XB_PITCH)),
                     "r"(nr * 3u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may go once it has been read
    }
}

cudaError_t rotate_orth(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int angle, cudaStream_t s)
{
    if (!w || !h) return cudaSuccess;
    if (angle == 180) {
        size_t npix = (size_t)w * h;
        if (npix % 16 == 0 && aligned16(src) && aligned16(dst)) {
            size_t n = npix / 16;  // one long row of npix pixels, mirrored
            if (n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
            launch(fliph_rgb16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s,
                   reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), (uint32_t)n, n);
        } else if (PPMX_VARIANT == 0 && npix >= 4096 && h <= 65535u * 64u) {  // any alignment: the tile kernel (ppmx_fused.cu)
            GeomOp go = {};
            go.rev_x = go.rev_y = 1;
            return geom_point(src, dst, w, h, 0, go, s);
        } else if (h <= 65535u && PPMX_VARIANT != 1) {  // any layout: mirrored rows taken from the mirrored row
            const unsigned gx = (unsigned)(((size_t)w * 3 / 4 + 1 + 255) / 256);
            launch(flip_rows_kernel<true, 3, true>, dim3(gx < 64u ? gx : 64u, h), dim3(256), 0, s, src, dst, w, h);
        } else {
            launch(reverse_pixels_kernel, dim3(wave_grid(npix * 3, 256, 8)), dim3(256), 0, s, src, dst, npix);
        }
        return PPMX_LAUNCHED();
    }
    if (angle != 90 && angle != 270) return cudaErrorInvalidValue;
    if ((w % 16u) == 0 && (h % 16u) == 0 && aligned16(src) && aligned16(dst) && (PPMX_VARIANT == 0 || PPMX_VARIANT == 9)) {
        dim3 grid((w + 63) / 64, (h + 63) / 64);
        if (grid.y > 65535u) return cudaErrorInvalidValue;
        if (angle == 90) launch(rotate_bulk64_kernel<true>, grid, dim3(128), 0, s, src, dst, w, h);
        else launch(rotate_bulk64_kernel<false>, grid, dim3(128), 0, s, src, dst, w, h);
        return PPMX_LAUNCHED();
    }
#ifdef PPMX_TUNING  // the register-path transposer the bulk-copy kernel replaced (variants 6, 7, 8 and its 2-tile default)
    if ((w % 16u) == 0 && (h % 16u) == 0 && aligned16(src) && aligned16(dst) && PPMX_VARIANT != 1) {
        // (numbering the CTAs down bands of 2..16 tile rows, for DRAM page locality on the write side,
        // measured 1-5 % SLOWER than this plain 2-D grid: the index arithmetic costs more than it gains)
        // two tiles per CTA at <= 40 registers (6 CTAs per SM) measured best: 0.725 / 0.828 of the HBM
        // roofline at 4096^2 / 16384^2; one tile per CTA (variant 6): 0.723 / 0.753
        const int nt = (PPMX_VARIANT == 6) ? 1 : 2;
        dim3 g64((w + XT * nt - 1) / (XT * nt), (h + XT - 1) / XT);
        if (g64.y > 65535u) return cudaErrorInvalidValue;
        if (PPMX_VARIANT == 6) {
            if (angle == 90) launch(rotate_transpose64_kernel<true, 1, 1, 0>, dim3(g64), dim3(256), 0, s, src, dst, w, h);
            else launch(rotate_transpose64_kernel<false, 1, 1, 0>, dim3(g64), dim3(256), 0, s, src, dst, w, h);
        } else if (PPMX_VARIANT == 7 || PPMX_VARIANT == 8) {
            const unsigned band = PPMX_VARIANT == 7 ? 8u : 4u;
            dim3 gb(g64.x * band, (g64.y + band - 1) / band);
            if (PPMX_VARIANT == 7) {
                if (angle == 90) launch(rotate_transpose64_kernel<true, 2, 6, 8>, gb, dim3(256), 0, s, src, dst, w, h);
                else launch(rotate_transpose64_kernel<false, 2, 6, 8>, gb, dim3(256), 0, s, src, dst, w, h);
            } else {
                if (angle == 90) launch(rotate_transpose64_kernel<true, 2, 6, 4>, gb, dim3(256), 0, s, src, dst, w, h);
                else launch(rotate_transpose64_kernel<false, 2, 6, 4>, gb, dim3(256), 0, s, src, dst, w, h);
            }
        } else {
            if (angle == 90) launch(rotate_transpose64_kernel<true, 2, 6, 0>, dim3(g64), dim3(256), 0, s, src, dst, w, h);
            else launch(rotate_transpose64_kernel<false, 2, 6, 0>, dim3(g64), dim3(256), 0, s, src, dst, w, h);
        }
        return PPMX_LAUNCHED();
    }
#endif
    if (PPMX_VARIANT == 0) {
        // any size, any alignment (1920 x 1080 frames: height no multiple of 16): the tile kernel of ppmx_fused.cu, rows in
        // by bulk copy or aligned words, out as 8-byte vectors or shifted words
        GeomOp go = {};
        go.transpose = 1;
        go.rev_x = angle == 90;   // out[x][h-1-y] = in[y][x], ref:717
        go.rev_y = angle == 270;  // out[w-1-x][y] = in[y][x], ref:725
        return geom_point(src, dst, w, h, 0, go, s);
    }
#ifdef PPMX_TUNING  // the byte-wise transposers the tile kernel replaced
    if (PPMX_VARIANT != 1) {
        dim3 ga((w + RA - 1) / RA, (h + RA - 1) / RA);
        if (ga.y > 65535u) return cudaErrorInvalidValue;
        if (angle == 90) launch(rotate_transpose_any_kernel<true>, ga, dim3(256), 0, s, src, dst, w, h);
        else launch(rotate_transpose_any_kernel<false>, ga, dim3(256), 0, s, src, dst, w, h);
        return PPMX_LAUNCHED();
    }
    dim3 grid((w + RT - 1) / RT, (h + RT - 1) / RT);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
    if (angle == 90) launch(rotate_transpose_kernel<true>, dim3(grid), dim3(256), 0, s, src, dst, w, h);
    else if (angle == 270) launch(rotate_transpose_kernel<false>, dim3(grid), dim3(256), 0, s, src, dst, w, h);
    else return cudaErrorInvalidValue;
    return PPMX_LAUNCHED();
}

#else
    return cudaErrorInvalidValue;  // (not reached: the tile kernel above takes every raster)
}
#endif

}  // namespace ppmx
