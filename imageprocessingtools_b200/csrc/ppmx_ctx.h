// ppmx_ctx.h -- internal: what ppmx_gpu.cu (contexts, rasters, single operators) and ppmx_chain.cu (op chains, row
// parts, several devices) share.  Not part of the public boundary (include/ppmx_gpu.h).
#pragma once

#include "../../include/ppmx_gpu.h"
#include "ppmx_kernels.h"

#include <cstdio>
#include <vector>

namespace ppmx {

constexpr int kLanes = 3;     // upload / kernels / download of consecutive rasters (or row parts of one) overlap
constexpr int kInFlight = 4;  // parts enqueued ahead of the GPU per lane: bounds what the pool holds

// the reference reports failures with one printf line on stdout (CHECK_ERROR, ref:31-36)
int fail(const char *what, cudaError_t e = cudaSuccess);

#define PPMX_CK(call, what)                                    \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return ::ppmx::fail(what, e__); \
    } while (0)

struct DeviceTables {  // imresize tables of one op, resident in HBM
    double *weights = nullptr;
    int *indices = nullptr;
    void *base = nullptr;
};

}  // namespace ppmx

struct ppmx_gpu_image {
    uint8_t *d = nullptr;
    uint32_t w = 0, h = 0;
    int layout = PPMX_LAYOUT_RGB8;
    size_t bytes = 0;
    int lane = 0;
};

// One CUDA device: streams ("lanes"), a PRIVATE stream-ordered pool (nothing here touches the device's default pool,
// which other libraries in the process may use), histogram scratch.  A context made by ppmx_gpu_init_multi owns one
// such context per device in `children` and has no streams of its own; single-device calls go to children[0].
struct ppmx_gpu_ctx {
    int device = 0;
    cudaStream_t lane[ppmx::kLanes] = {};
    cudaEvent_t tables_ready = nullptr;
    cudaEvent_t lane_done[ppmx::kLanes] = {};
    cudaMemPool_t pool = nullptr;
    unsigned long long *d_hist = nullptr;  // 256 bins
    unsigned long long *h_hist = nullptr;  // pinned copy
    std::vector<ppmx_gpu_ctx *> children;  // multi-device context only
};

namespace ppmx {

inline ppmx_gpu_ctx *primary(ppmx_gpu_ctx *c) { return (c && !c->children.empty()) ? c->children[0] : c; }

// stream-ordered allocation from the context's private pool
cudaError_t pool_alloc(ppmx_gpu_ctx *c, void **p, size_t bytes, cudaStream_t s);

int image_alloc_on(ppmx_gpu_ctx *c, int lane, uint32_t w, uint32_t h, int layout, ppmx_gpu_image **out);
void image_free_on(ppmx_gpu_ctx *c, ppmx_gpu_image *im);
int upload_tables(ppmx_gpu_ctx *c, const ppmx_op *op, DeviceTables *t, cudaStream_t s);

// launches the kernel(s) of one operator on raw device pointers
int launch_op(const ppmx_op *op, const uint8_t *d_src, uint32_t w, uint32_t h, int layout, uint8_t *d_dst, const Band &band,
              unsigned long long *d_hist, const DeviceTables *tables, cudaStream_t s);

inline size_t row_bytes(uint32_t w, int layout)
{
    return layout == PPMX_LAYOUT_RGB8 ? (size_t)w * 3 : layout == PPMX_LAYOUT_R8 ? (size_t)w : (size_t)((w + 7u) / 8u);
}

}  // namespace ppmx
