// ppmx_conv_vw.cu -- EXTENSION (no reference counterpart, parity unpinned): dense k x k integer convolution for k = 9 .. 15 (any
// signed-byte coefficients) on vertical words kept in shared memory.  Part of libppmx_gpu.so; conventions in ppmx_common.cuh.
//
// The 5x5 / 7x7 strip kernels keep a thread's 8-row window in registers and fetch the neighbours' columns by shuffle; from
// k = 9 on neither fits (window of 16 rows, 3 (k/2) > 16 halo columns), and everything used to fall to the scalar kernel
// (one multiply-add per tap per byte from shared-memory bytes: 0.003-0.01 of the HBM roofline).  Here a CTA stages a tile of
// (32 + k - 1) source rows x 192 byte columns as VERTICAL WORDS -- word (g, c) = rows 4g .. 4g+3 of byte column c, made by 4x4
// byte transposes on the way in -- and an output (row t, column c) is, per tap column j, the dot products of the words that hold
// rows t .. t+k-1 of column c + 3 (j - k/2) against that tap column's coefficients shifted to the row phase t % 4:
// k * ceil((k + phase) / 4) dp4a per byte (27 at 9x9) instead of k^2 scalar multiply-adds.  A warp owns one group of four output
// rows (its phases are compile-time constants), a thread four consecutive byte columns of them: 16 accumulators, coalesced
// 4-byte stores.  Columns are laid out residue-major (column c at word (c % 4) * 48 + c / 4) so that the 32 lanes of a warp, which
// read columns 4 apart, hit consecutive banks.
#include "ppmx_conv.cuh"

namespace ppmx {

constexpr int VW_TW = 128;    // byte columns of outputs per tile
constexpr int VW_TH = 32;     // output rows per tile: 8 warps x 4 rows
constexpr int VW_HALO = 32;   // staged byte columns either side (>= 3 * 7 + 3, a multiple of 16)
constexpr int VW_NCOL = VW_TW + 2 * VW_HALO;  // 192 staged byte columns = 12 vectors per row

template <int K>
struct VwCoef {
    static constexpr int NWD = (K + 3 + 3) / 4;  // words spanned by K rows starting at phase 3
    uint32_t cw[4][K][NWD];                      // [row phase][tap column][word]: coefficient bytes in the rows they multiply
};

__device__ __forceinline__ int vw_slot(int col) { return (col & 3) * (VW_NCOL / 4) + (col >> 2); }

// stage the tile's source rows as vertical words: item = (row group g, vector v): four rows x 16 bytes -> 16 words.
// PLANAR_COLS = false: residue-major columns (vw_slot) for the dense kernel; true: columns in order for the rank-1 kernel.
template <int NG, bool IN_ORDER, int NCOL, int HALO>
__device__ __forceinline__ void vw_stage(const RowSource &rs, uint32_t (*vw)[NCOL], const uint8_t **rowp, int x0, int ys, int R,
                                         uint32_t row_bytes, int tid)
{
    const size_t pitch = row_bytes;
    const int w = (int)(row_bytes / 3u);
    // the 4 NG source rows of the tile, resolved once (mirror at the raster's top / bottom, halo rows of a band)
    if (tid < 4 * NG) rowp[tid] = rs.row(rs.y0 + ys - R + tid, pitch);
    __syncthreads();
    for (int item = tid; item < NG * (NCOL / 16); item += 256) {
        const int g = item / (NCOL / 16), v = item - g * (NCOL / 16);
        const int c0 = x0 - HALO + 16 * v;  // first byte column of the vector (may lie outside the row)
        uint32_t rw[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint8_t *row = rowp[4 * g + i];
            if (c0 >= 0 && c0 + 16 <= (int)row_bytes) {
                const uint4 q = *reinterpret_cast<const uint4 *>(row + c0);
                rw[i][0] = q.x, rw[i][1] = q.y, rw[i][2] = q.z, rw[i][3] = q.w;
            } else {  // the raster's left / right edge: mirrored pixels, byte by byte
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t word = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int c = c0 + 4 * q + b;
                        const int px = c >= 0 ? c / 3 : -((-c + 2) / 3), ch = c - 3 * px;  // floor division: ch in 0..2
                        word |= (uint32_t)row[(size_t)mirror_index(px, w) * 3 + ch] << (8 * b);
                    }
                    rw[i][q] = word;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {  // 4x4 byte transpose of the four rows' word q -> columns 16v + 4q .. + 3
            const uint32_t t0 = __byte_perm(rw[0][q], rw[1][q], 0x5140), t1 = __byte_perm(rw[2][q], rw[3][q], 0x5140);
            const uint32_t t2 = __byte_perm(rw[0][q], rw[1][q], 0x7362), t3 = __byte_perm(rw[2][q], rw[3][q], 0x7362);
            const uint32_t c0w = __byte_perm(t0, t1, 0x5410), c1w = __byte_perm(t0, t1, 0x7632), c2w = __byte_perm(t2, t3, 0x5410),
                           c3w = __byte_perm(t2, t3, 0x7632);
            if (IN_ORDER) {
                *reinterpret_cast<uint4 *>(&vw[g][16 * v + 4 * q]) = make_uint4(c0w, c1w, c2w, c3w);
            } else {
                const int base = 4 * v + q;  // (column >> 2); the column's residue = the word's index here
                vw[g][0 * (NCOL / 4) + base] = c0w;
                vw[g][1 * (NCOL / 4) + base] = c1w;
                vw[g][2 * (NCOL / 4) + base] = c2w;
                vw[g][3 * (NCOL / 4) + base] = c3w;
            }
        }
    }
}

template <int K, int MODE>
__global__ void __launch_bounds__(256) conv_vw_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t row_bytes, const VwCoef<K> cf,
                                                      const ConvRound rnd)
{
    constexpr int R = K / 2, NRS = VW_TH + K - 1, NG = (NRS + 3) / 4, NWD = VwCoef<K>::NWD;
    __shared__ __align__(16) uint32_t vw[NG][VW_NCOL];
    __shared__ const uint8_t *rowp[4 * NG];
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * VW_TW, ys = blockIdx.y * VW_TH;  // first output byte column / row (band-local) of the tile
    const size_t pitch = row_bytes;
    pdl_wait();

    vw_stage<NG, false, VW_NCOL, VW_HALO>(rs, vw, rowp, x0, ys, R, row_bytes, tid);
    __syncthreads();

    // ---- compute: warp = output rows 4 warp .. 4 warp + 3 of the tile, thread = byte columns 4 lane .. 4 lane + 3 ----
    const int cx = x0 + 4 * lane;
    if (cx >= (int)row_bytes || ys + 4 * warp >= rs.h) return;
    int32_t acc[4][4];  // [row phase][column]
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int d = 0; d < 4; d++) acc[p][d] = rnd.start;
    // staged column of output column d under tap column j: VW_HALO + 4 lane + d + 3 (j - R); all columns a thread touches are
    // VW_HALO + 4 lane + e with e = d + 3 (j - R) in -3R .. 3R + 3: walk e, feed every (d, j) that lands on it
#pragma unroll
    for (int e = -3 * R; e <= 3 * R + 3; e++) {
        const int col = VW_HALO + e;  // + 4 lane below: (col + 4 lane) & 3 == col & 3, (col + 4 lane) >> 2 == (col >> 2) + lane
        const uint32_t *p = &vw[warp][(col & 3) * (VW_NCOL / 4) + (col >> 2) + lane];
        uint32_t wd[NWD];
#pragma unroll
        for (int q = 0; q < NWD; q++) wd[q] = p[q * VW_NCOL];  // row groups warp .. warp + NWD - 1 (staged row 0 = output row 0 - R)
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const int j3 = e - d;  // = 3 (j - R)
            if (j3 % 3 != 0) continue;
            const int j = j3 / 3 + R;
            if (j < 0 || j >= K) continue;
#pragma unroll
            for (int ph = 0; ph < 4; ph++)
#pragma unroll
                for (int q = 0; q < NWD; q++)
                    if (4 * q < ph + K && 4 * q + 3 >= ph)  // word q holds one of the rows ph .. ph + K - 1
                        acc[ph][d] = dp4a_u8s8(wd[q], cf.cw[ph][j][q], acc[ph][d]);
        }
    }
#pragma unroll
    for (int ph = 0; ph < 4; ph++) {
        const int y = ys + 4 * warp + ph;
        if (y >= rs.h) break;
        *reinterpret_cast<uint32_t *>(dst + (size_t)y * pitch + cx) = rnd.template pack4<MODE>(acc[ph][0], acc[ph][1], acc[ph][2], acc[ph][3]);
    }
}

template <int K>
static bool conv_vw_k(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef, const ConvRound &rnd, cudaStream_t s,
                      cudaError_t *err)
{
    constexpr int NWD = VwCoef<K>::NWD;
    VwCoef<K> cf;
    for (int ph = 0; ph < 4; ph++)
        for (int j = 0; j < K; j++) {
            for (int q = 0; q < NWD; q++) cf.cw[ph][j][q] = 0;
            for (int i = 0; i < K; i++) {  // tap row i multiplies staged row ph + i of the warp's first group
                const int32_t c = coef[i * K + j];
                if (c < -128 || c > 127) return false;
                cf.cw[ph][j][(ph + i) >> 2] |= (uint32_t)(uint8_t)(int8_t)c << (8 * ((ph + i) & 3));
            }
        }
    const uint32_t row_bytes = w * 3u;
    dim3 grid((row_bytes + VW_TW - 1) / VW_TW, (h + VW_TH - 1) / VW_TH);
    if (grid.y > 65535u) {
        *err = cudaErrorInvalidValue;
        return true;
    }
    if (rnd.mode == 0) launch(conv_vw_kernel<K, 0>, grid, dim3(256), 0, s, rs, dst, row_bytes, cf, rnd);
    else if (rnd.mode == 1) launch(conv_vw_kernel<K, 1>, grid, dim3(256), 0, s, rs, dst, row_bytes, cf, rnd);
    else launch(conv_vw_kernel<K, 2>, grid, dim3(256), 0, s, rs, dst, row_bytes, cf, rnd);
    *err = PPMX_LAUNCHED();
    return true;
}

bool conv_vw(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, const ConvRound &rnd, cudaStream_t s,
             cudaError_t *err)
{
    if (k == 9) return conv_vw_k<9>(rs, dst, w, h, coef, rnd, s, err);
    if (k == 11) return conv_vw_k<11>(rs, dst, w, h, coef, rnd, s, err);
    if (k == 13) return conv_vw_k<13>(rs, dst, w, h, coef, rnd, s, err);
    if (k == 15) return conv_vw_k<15>(rs, dst, w, h, coef, rnd, s, err);
    return false;
}

// ---- rank-1 kernels (coef = u v^T, non-negative, column sums within 16 bits: Gaussian / binomial blurs), k = 9 .. 15 -------------
// The same staged words, columns in order.  Pass A: every thread computes column sums S (rows t .. t+k-1 of a byte column against
// u: ceil((k + phase) / 4) dp4a) for the tile's 32 rows x 160 columns and stores them as 16-bit values, "residue-planar": column c in
// plane c % 3 at entry c / 3 of its row, so that horizontal neighbours of one channel (columns 3 apart) are adjacent halves.
// Pass B: an output is (k + 1) / 2 dp2a over consecutive words of that row -- taps pair up from an even position; an output whose
// first tap sits at an odd position uses the coefficient bytes moved up one place.
template <int K>
struct VwSepCoef {
    static constexpr int NWD = (K + 3 + 3) / 4, NPW = (K + 1) / 2, NVW = (K + 1 + 3) / 4;
    uint32_t ucw[4][NWD];  // [row phase][word]: u in the bytes of the rows it multiplies
    uint32_t va[NVW];      // bytes v0 .. v(K-1), 0     (first tap at an even position)
    uint32_t vb[NVW];      // bytes 0, v0 .. v(K-1)     (first tap at an odd position)
};

constexpr int VS_TW = 96;     // byte columns of outputs per tile = 32 pixels: lane l owns pixel l
constexpr int VS_HALO = 32;   // staged byte columns either side (keeps the vectors 16-byte aligned: 96 = 6 x 16)
constexpr int VS_NCOL = VS_TW + 2 * VS_HALO;

template <int K, int MODE>
__global__ void __launch_bounds__(256) conv_vwsep_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t row_bytes,
                                                         const VwSepCoef<K> cf, const ConvRound rnd)
{
    constexpr int R = K / 2, NRS = VW_TH + K - 1, NG = (NRS + 3) / 4, NWD = VwSepCoef<K>::NWD, NPW = VwSepCoef<K>::NPW;
    constexpr int NVW = VwSepCoef<K>::NVW;
    __shared__ __align__(16) uint32_t vw[NG][VS_NCOL];
    // column sums, 16 bits each: per row three residue planes, 43 words apart (32 would put the planes on the same banks)
    constexpr int PLANE = 43, SROW = 3 * PLANE;
    __shared__ __align__(16) uint32_t s32[VW_TH][SROW];
    __shared__ const uint8_t *rowp[4 * NG];
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * VS_TW, ys = blockIdx.y * VW_TH;
    const size_t pitch = row_bytes;
    pdl_wait();
    vw_stage<NG, true, VS_NCOL, VS_HALO>(rs, vw, rowp, x0, ys, R, row_bytes, tid);
    __syncthreads();

    // ---- pass A: warp = tile rows 4 warp .. 4 warp + 3, lane = staged columns lane, lane + 32, ... ----
#pragma unroll
    for (int m = 0; m < VS_NCOL / 32; m++) {
        const int col = lane + 32 * m;
        uint32_t wd[NWD];
#pragma unroll
        for (int q = 0; q < NWD; q++) wd[q] = vw[warp + q][col];
        const int px = col / 3, res = col - 3 * px;
#pragma unroll
        for (int ph = 0; ph < 4; ph++) {
            uint32_t sum = 0;
#pragma unroll
            for (int q = 0; q < NWD; q++)
                if (4 * q < ph + K && 4 * q + 3 >= ph) sum = __dp4a(wd[q], cf.ucw[ph][q], sum);
            reinterpret_cast<uint16_t *>(s32[4 * warp + ph])[res * (2 * PLANE) + px] = (uint16_t)sum;
        }
    }
    __syncthreads();

    // ---- pass B: warp = the same four rows, lane = pixel `lane` of the tile: byte columns 3 lane + r, staged column
    // 32 + 3 lane + r = 3 (lane + 10) + 2 | 3 (lane + 11) | 3 (lane + 11) + 1: residue and parity are per-r constants up to `lane`,
    // consecutive lanes read consecutive half-words (two lanes per word: no bank conflict) ----
    if (ys + 4 * warp >= rs.h) return;  // (a whole warp)
    uint32_t word_off[3], cv[3][NVW];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int res = r == 0 ? 2 : r - 1, start = lane + (r == 0 ? 10 : 11) - R;  // first tap's position in its plane
        word_off[r] = (uint32_t)(res * PLANE + (start >> 1));
#pragma unroll
        for (int q = 0; q < NVW; q++) cv[r][q] = (start & 1) ? cf.vb[q] : cf.va[q];  // taps pair up from an even position
    }
    const int sub = lane & 3, word = 3 * (lane >> 2) + sub;  // lanes 4q .. 4q+2 assemble the three words of pixels 4q .. 4q+3
    const bool stores = sub < 3 && x0 + 4 * word < (int)row_bytes;
#pragma unroll
    for (int ph = 0; ph < 4; ph++) {
        const int y = ys + 4 * warp + ph;
        if (y >= rs.h) break;
        const uint32_t *srow = s32[4 * warp + ph];
        int32_t a[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const uint32_t *p = srow + word_off[r];
            uint32_t t = (uint32_t)rnd.start;
#pragma unroll
            for (int m = 0; m < NPW; m++) {
                const uint32_t cw = cv[r][m >> 1], q = p[m];
                if (m & 1) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(q), "r"(cw), "r"(t));
                else asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(q), "r"(cw), "r"(t));
            }
            a[r] = (int32_t)t;
        }
        // the pixel's three bytes, then four pixels' 12 bytes as three words
        const uint32_t T = rnd.template pack4<MODE>(a[0], a[1], a[2], a[2]) & 0x00ffffffu;
        const uint32_t Tn = __shfl_down_sync(0xffffffffu, T, 1);
        const uint32_t out = __funnelshift_r(T | (Tn << 24), Tn >> 8, 8u * (uint32_t)sub);
        if (stores) *reinterpret_cast<uint32_t *>(dst + (size_t)y * pitch + x0 + 4 * word) = out;
    }
}

template <int K>
static bool conv_vwsep_k(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef, int32_t div, int32_t bias,
                         ConvRound rnd, cudaStream_t s, cudaError_t *err)
{
    constexpr int NWD = VwSepCoef<K>::NWD, NVW = VwSepCoef<K>::NVW;
    int32_t u[K], v[K];
    if (!rank_one<K>(coef, u, v)) return false;
    bool neg = true, pos = true;
    int64_t su = 0, sv = 0;
    for (int i = 0; i < K; i++) neg = neg && u[i] <= 0 && v[i] <= 0, pos = pos && u[i] >= 0 && v[i] >= 0;
    if (neg)
        for (int i = 0; i < K; i++) u[i] = -u[i], v[i] = -v[i];
    else if (!pos) return false;
    for (int i = 0; i < K; i++) su += u[i], sv += v[i];
    if (255 * su > 65535) return false;
    for (int i = 0; i < K; i++)
        if (u[i] > 255 || v[i] > 255) return false;
    int mode = 2;
    if (rnd.mode == 1 && bias == 0 && div <= 65536 && (div & (div - 1)) == 0 && su * sv == div) {  // the quotient as byte 2 of the sum
        const int64_t scale = 65536 / div;
        for (int64_t a = scale; a >= 1 && mode != 4; a >>= 1) {
            const int64_t b = scale / a;
            bool ok = 255 * su * a <= 65535;
            for (int i = 0; i < K; i++) ok = ok && u[i] * a <= 255 && v[i] * b <= 255;
            if (ok) {
                for (int i = 0; i < K; i++) u[i] = (int32_t)(u[i] * a), v[i] = (int32_t)(v[i] * b);
                rnd.start = 32768;
                mode = 4;
            }
        }
    }
    if (mode == 2) rnd.mode = 2, rnd.start = 0;  // (the multiply-high constants are filled in whatever mode conv() chose)
    VwSepCoef<K> cf;
    for (int ph = 0; ph < 4; ph++) {
        for (int q = 0; q < NWD; q++) cf.ucw[ph][q] = 0;
        for (int i = 0; i < K; i++) cf.ucw[ph][(ph + i) >> 2] |= (uint32_t)u[i] << (8 * ((ph + i) & 3));
    }
    for (int q = 0; q < NVW; q++) cf.va[q] = cf.vb[q] = 0;
    for (int i = 0; i < K; i++) {
        cf.va[i >> 2] |= (uint32_t)v[i] << (8 * (i & 3));
        cf.vb[(i + 1) >> 2] |= (uint32_t)v[i] << (8 * ((i + 1) & 3));
    }
    const uint32_t row_bytes = w * 3u;
    dim3 grid((row_bytes + VS_TW - 1) / VS_TW, (h + VW_TH - 1) / VW_TH);
    if (grid.y > 65535u) {
        *err = cudaErrorInvalidValue;
        return true;
    }
    if (mode == 4) launch(conv_vwsep_kernel<K, 4>, grid, dim3(256), 0, s, rs, dst, row_bytes, cf, rnd);
    else launch(conv_vwsep_kernel<K, 2>, grid, dim3(256), 0, s, rs, dst, row_bytes, cf, rnd);
    *err = PPMX_LAUNCHED();
    return true;
}

bool conv_vwsep(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div, int32_t bias,
                const ConvRound &rnd, cudaStream_t s, cudaError_t *err)
{
    if (k == 9) return conv_vwsep_k<9>(rs, dst, w, h, coef, div, bias, rnd, s, err);
    if (k == 11) return conv_vwsep_k<11>(rs, dst, w, h, coef, div, bias, rnd, s, err);
    if (k == 13) return conv_vwsep_k<13>(rs, dst, w, h, coef, div, bias, rnd, s, err);
    if (k == 15) return conv_vwsep_k<15>(rs, dst, w, h, coef, div, bias, rnd, s, err);
    return false;
}

}  // namespace ppmx
