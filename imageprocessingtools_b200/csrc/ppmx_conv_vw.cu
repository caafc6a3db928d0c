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

template <int K, int MODE>
__global__ void __launch_bounds__(256) conv_vw_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t row_bytes, const VwCoef<K> cf,
                                                      const ConvRound rnd)
{
    constexpr int R = K / 2, NRS = VW_TH + K - 1, NG = (NRS + 3) / 4, NWD = VwCoef<K>::NWD;
    __shared__ __align__(16) uint32_t vw[NG][VW_NCOL];
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * VW_TW, ys = blockIdx.y * VW_TH;  // first output byte column / row (band-local) of the tile
    const size_t pitch = row_bytes;
    const int w = (int)(row_bytes / 3u);
    pdl_wait();

    // ---- stage: item = (row group g, vector v): four rows x 16 bytes -> 16 vertical words ----
    for (int item = tid; item < NG * (VW_NCOL / 16); item += 256) {
        const int g = item / (VW_NCOL / 16), v = item - g * (VW_NCOL / 16);
        const int c0 = x0 - VW_HALO + 16 * v;  // first byte column of the vector (may lie outside the row)
        uint32_t rw[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint8_t *row = rs.row(rs.y0 + ys - R + 4 * g + i, pitch);  // mirror at the raster's top / bottom, halo rows of a band
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
            const int base = 4 * v + q;  // (column >> 2); residue r of the column = the word's index below
            vw[g][0 * (VW_NCOL / 4) + base] = __byte_perm(t0, t1, 0x5410);
            vw[g][1 * (VW_NCOL / 4) + base] = __byte_perm(t0, t1, 0x7632);
            vw[g][2 * (VW_NCOL / 4) + base] = __byte_perm(t2, t3, 0x5410);
            vw[g][3 * (VW_NCOL / 4) + base] = __byte_perm(t2, t3, 0x7632);
        }
    }
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

}  // namespace ppmx
