// ppmx_kernels.h -- internal launch interface between the C ABI (ppmx_gpu.cu) and the
// sm_100a kernels (ppmx_kernels.cu).  Not part of the public boundary (include/ppmx_gpu.h).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

namespace ppmx {

// Row-band context for operators whose result depends on the absolute row (Bayer phase,
// mirror border) or on rows owned by a neighbouring GPU (halo).  Zero-initialised = whole image.
struct Band {
    uint32_t full_h = 0;           // height of the whole raster (0: the band IS the raster)
    uint32_t y0 = 0;               // first row of this band within the whole raster
    const uint8_t *top = nullptr;  // `halo` rows directly above the band (may be peer memory)
    const uint8_t *bottom = nullptr;
    uint32_t halo = 0;
    uint32_t out_y0 = 0, out_rows = 0;  // imresize height pass: the slice of output rows this band produces
};

// Which implementation of an operator to launch; 0 = the default (best measured).  Only the tuning build
// (-DPPMX_TUNING, libppmx_gpu_tuning.so) has the switch: in the release library PPMX_VARIANT is the constant 0, so the
// alternative kernels are never instantiated and do not ship.
#ifdef PPMX_TUNING
extern int g_variant;
#define PPMX_VARIANT (::ppmx::g_variant)
#else
#define PPMX_VARIANT 0
#endif
extern std::atomic<int> g_pdl;  // programmatic dependent launch on/off (process-wide)

cudaError_t gray(const uint8_t *src, uint8_t *dst, size_t npix, unsigned long long *d_hist, cudaStream_t s);
cudaError_t hist_gray(const uint8_t *src, size_t npix, unsigned long long *d_hist, cudaStream_t s);
cudaError_t mono_plane(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t y0, cudaStream_t s);
cudaError_t mono_bits(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t y0, cudaStream_t s);
cudaError_t pack_pbm(const uint8_t *src, int src_bpp, uint8_t *dst, uint32_t w, uint32_t h, cudaStream_t s);
cudaError_t extract_r(const uint8_t *src, uint8_t *dst, size_t npix, cudaStream_t s);
cudaError_t flip(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int bpp, int vertical, cudaStream_t s);
cudaError_t rotate_orth(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int angle, cudaStream_t s);
cudaError_t rotate_bicubic(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh,
                           double cos_t, double sin_t, cudaStream_t s);
// one separable pass; d_weights/d_indices are [out_size][taps] in device memory
cudaError_t imresize(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int out_size, int dim, int taps,
                     const double *d_weights, const int *d_indices, const Band &band, cudaStream_t s);
// extension operators (no reference counterpart)
cudaError_t conv(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef /*host*/,
                 int32_t div, int32_t bias, const Band &band, cudaStream_t s);

// Geometry + pointwise tail in one pass (ppmx_fused.cu).  Output pixel (X, Y) = point(source pixel (x, y)):
//   transpose 0: x = rev_x ? w-1-X : X,  y = rev_y ? h-1-Y : Y          (flips, 180 degrees)
//   transpose 1: y = rev_x ? h-1-X : X,  x = rev_y ? w-1-Y : Y          (90 / 270 degrees, with or without a flip)
// point: 0 RGB8 -> RGB8, 1 grey -> R8 (ref:1000), 2 .r -> R8 (ref:263-267), 3 mono -> packed bits (ref:964-969 + 268-284),
// whose Bayer index (xm%4)*4 + (ym%4) is taken at xm = +-(mx_from_y ? y : x) + mx_add, ym = +-(mx_from_y ? x : y) + my_add.
struct GeomOp {
    int transpose, rev_x, rev_y, point;
    int mx_from_y, mx_neg, mx_add, my_neg, my_add;
};
bool geom_point_supported(uint32_t w, uint32_t h, const GeomOp &go);
// in_pitch: bytes between source rows (0 = 3 * w; larger for a column slab of a wider raster)
cudaError_t geom_point(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t in_pitch, const GeomOp &go,
                       cudaStream_t s);

cudaError_t geom_repitch(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t in_pitch, uint32_t out_pitch,
                         cudaStream_t s);

// every byte of the raster through a 256-entry table (host pointer)
cudaError_t levels(const uint8_t *src, uint8_t *dst, size_t nbytes, const uint8_t *lut /*host*/, cudaStream_t s);

unsigned long long launch_count();
void add_launches(unsigned long long n);  // kernels replayed by a CUDA graph launch

}  // namespace ppmx
