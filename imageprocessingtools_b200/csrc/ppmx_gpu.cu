// ppmx_gpu.cu -- implementation of the C ABI declared in include/ppmx_gpu.h.
//
// Owns: one CUDA device per context, a small set of streams ("lanes") with stream-ordered
// device allocations, the buff/new_buff hand-over rules of the reference's op chain
// (ref:1084-1155 = /root/reference/ppmx-edward.c), and the pinned upload/download path.
// All arithmetic lives in ppmx_kernels.cu.  There is no CPU fallback anywhere in this file.
#include "../../include/ppmx_gpu.h"
#include "ppmx_kernels.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using ppmx::Band;

#define PPMX_VERSION "ppmx-b200 0.1 (sm_100a)"

namespace {

constexpr int kLanes = 3;  // upload / compute / download of consecutive rasters overlap

// the reference reports failures with one printf line on stdout (CHECK_ERROR, ref:31-36)
int fail(const char *what, cudaError_t e = cudaSuccess)
{
    if (e != cudaSuccess) printf("ppmx_gpu error: %s: %s\n", what, cudaGetErrorString(e));
    else printf("ppmx_gpu error: %s\n", what);
    fflush(stdout);
    return PPMX_ERROR;
}

#define CK(call, what)                                         \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return fail(what, e__);        \
    } while (0)

struct DeviceTables {  // imresize tables of one op, resident in HBM
    double *weights = nullptr;
    int *indices = nullptr;
    void *base = nullptr;
};

}  // namespace

struct ppmx_gpu_image {
    uint8_t *d = nullptr;
    uint32_t w = 0, h = 0;
    int layout = PPMX_LAYOUT_RGB8;
    size_t bytes = 0;
    int lane = 0;
};

struct ppmx_gpu_ctx {
    int device = 0;
    cudaStream_t lane[kLanes] = {};
    cudaEvent_t tables_ready = nullptr;
    unsigned long long *d_hist = nullptr;   // 256 bins
    unsigned long long *h_hist = nullptr;   // pinned copy
};

extern "C" size_t ppmx_gpu_layout_bytes(uint32_t w, uint32_t h, int layout)
{
    switch (layout) {
    case PPMX_LAYOUT_RGB8: return (size_t)w * h * 3;
    case PPMX_LAYOUT_R8: return (size_t)w * h;
    case PPMX_LAYOUT_BITS: return (size_t)((w + 7u) / 8u) * h;
    default: return 0;
    }
}

extern "C" int ppmx_gpu_set_tuning(const char *key, int value)
{
    if (key && !strcmp(key, "variant")) ppmx::g_variant = value;
    else if (key && !strcmp(key, "pdl")) ppmx::g_pdl = value ? 1 : 0;
    else return PPMX_ERROR;
    return PPMX_OK;
}

extern "C" const char *ppmx_gpu_version(void) { return PPMX_VERSION; }
extern "C" uint64_t ppmx_gpu_launch_count(void) { return ppmx::launch_count(); }

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------

extern "C" int ppmx_gpu_init(ppmx_gpu_ctx **out, int device)
{
    if (!out) return fail("ppmx_gpu_init: null ctx pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail("no CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= n) return fail("ppmx_gpu_init: no such device");
    CK(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10) return fail("this build carries sm_100a code only and needs a B200-class device");

    ppmx_gpu_ctx *c = new (std::nothrow) ppmx_gpu_ctx();
    if (!c) return fail("out of host memory");
    c->device = device;
    for (int i = 0; i < kLanes; i++) CK(cudaStreamCreateWithFlags(&c->lane[i], cudaStreamNonBlocking), "cudaStreamCreate");
    CK(cudaEventCreateWithFlags(&c->tables_ready, cudaEventDisableTiming), "cudaEventCreate");
    CK(cudaMalloc(&c->d_hist, 256 * sizeof(unsigned long long)), "cudaMalloc hist");
    CK(cudaMallocHost(&c->h_hist, 256 * sizeof(unsigned long long)), "cudaMallocHost hist");
    // keep freed rasters cached in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = c;
    return PPMX_OK;
}

extern "C" void ppmx_gpu_free(ppmx_gpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < kLanes; i++)
        if (c->lane[i]) {
            cudaStreamSynchronize(c->lane[i]);
            cudaStreamDestroy(c->lane[i]);
        }
    if (c->tables_ready) cudaEventDestroy(c->tables_ready);
    cudaMemPool_t pool;  // hand the cached rasters back to the driver
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    if (c->d_hist) cudaFree(c->d_hist);
    if (c->h_hist) cudaFreeHost(c->h_hist);
    delete c;
}

extern "C" void *ppmx_gpu_host_alloc(ppmx_gpu_ctx *c, size_t bytes)
{
    void *p = nullptr;
    if (c) cudaSetDevice(c->device);
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail("cudaMallocHost");
        return nullptr;
    }
    return p;
}

extern "C" void ppmx_gpu_host_free(ppmx_gpu_ctx *c, void *p)
{
    (void)c;
    if (p) cudaFreeHost(p);
}

extern "C" int ppmx_gpu_sync(ppmx_gpu_ctx *c)
{
    if (!c) return fail("null ctx");
    for (int i = 0; i < kLanes; i++) CK(cudaStreamSynchronize(c->lane[i]), "cudaStreamSynchronize");
    return PPMX_OK;
}

// ---------------------------------------------------------------------------------------------
// rasters in HBM
// ---------------------------------------------------------------------------------------------

static int image_alloc_on(ppmx_gpu_ctx *c, int lane, uint32_t w, uint32_t h, int layout, ppmx_gpu_image **out)
{
    size_t bytes = ppmx_gpu_layout_bytes(w, h, layout);
    ppmx_gpu_image *im = new (std::nothrow) ppmx_gpu_image();
    if (!im) return fail("out of host memory");
    im->w = w;
    im->h = h;
    im->layout = layout;
    im->bytes = bytes;
    im->lane = lane;
    // +16: vector kernels never read past `bytes`, the slack only keeps zero-sized rasters valid
    cudaError_t e = cudaMallocAsync((void **)&im->d, bytes + 16, c->lane[lane]);
    if (e != cudaSuccess) {
        delete im;
        return fail("can not allocate image buff in HBM", e);  // wording of ref:926
    }
    *out = im;
    return PPMX_OK;
}

static void image_free_on(ppmx_gpu_ctx *c, ppmx_gpu_image *im)
{
    if (!im) return;
    if (im->d) cudaFreeAsync(im->d, c->lane[im->lane]);
    delete im;
}

extern "C" int ppmx_gpu_image_alloc(ppmx_gpu_ctx *c, uint32_t w, uint32_t h, int layout, ppmx_gpu_image **img)
{
    if (!c || !img) return fail("ppmx_gpu_image_alloc: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    return image_alloc_on(c, 0, w, h, layout, img);
}

extern "C" void ppmx_gpu_image_free(ppmx_gpu_ctx *c, ppmx_gpu_image *img)
{
    if (!c) return;
    cudaSetDevice(c->device);
    image_free_on(c, img);
}

extern "C" int ppmx_gpu_image_info(const ppmx_gpu_image *img, uint32_t *w, uint32_t *h, int *layout, size_t *bytes,
                                   void **device_ptr)
{
    if (!img) return fail("ppmx_gpu_image_info: null image");
    if (w) *w = img->w;
    if (h) *h = img->h;
    if (layout) *layout = img->layout;
    if (bytes) *bytes = img->bytes;
    if (device_ptr) *device_ptr = img->d;
    return PPMX_OK;
}

static int upload_on(ppmx_gpu_ctx *c, int lane, const uint8_t *src, uint32_t w, uint32_t h, int layout,
                     ppmx_gpu_image **img)
{
    if (image_alloc_on(c, lane, w, h, layout, img) != PPMX_OK) return PPMX_ERROR;
    if ((*img)->bytes) CK(cudaMemcpyAsync((*img)->d, src, (*img)->bytes, cudaMemcpyHostToDevice, c->lane[lane]), "upload");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_upload(ppmx_gpu_ctx *c, const uint8_t *src, uint32_t w, uint32_t h, int layout,
                               ppmx_gpu_image **img)
{
    if (!c || !img || (!src && w && h)) return fail("ppmx_gpu_upload: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    return upload_on(c, 0, src, w, h, layout, img);
}

// ---------------------------------------------------------------------------------------------
// one operator
// ---------------------------------------------------------------------------------------------

extern "C" int ppmx_gpu_op_output(const ppmx_op *op, uint32_t w, uint32_t h, int layout, uint32_t *ow, uint32_t *oh,
                                  int *olayout)
{
    if (!op) return PPMX_ERROR;
    uint32_t nw = w, nh = h;
    int nl = layout;
    switch (op->kind) {
    case PPMX_OP_GRAY:
    case PPMX_OP_GRAY_HIST:
    case PPMX_OP_MONO:
    case PPMX_OP_EXTRACT_R:
        if (layout != PPMX_LAYOUT_RGB8) return PPMX_ERROR;
        nl = PPMX_LAYOUT_R8;
        break;
    case PPMX_OP_MONO_BITS:
        if (layout != PPMX_LAYOUT_RGB8) return PPMX_ERROR;
        nl = PPMX_LAYOUT_BITS;
        break;
    case PPMX_OP_PACK_PBM:
        if (layout == PPMX_LAYOUT_BITS) return PPMX_ERROR;
        nl = PPMX_LAYOUT_BITS;
        break;
    case PPMX_OP_FLIP:
        if (layout == PPMX_LAYOUT_BITS) return PPMX_ERROR;
        break;
    case PPMX_OP_ROTATE:
        if (layout != PPMX_LAYOUT_RGB8) return PPMX_ERROR;
        if (op->angle_deg != 0) {  // ref:701-705: angle 0 keeps the buffer and its size
            nw = op->new_width;
            nh = op->new_height;
        }
        break;
    case PPMX_OP_IMRESIZE:
        if (layout != PPMX_LAYOUT_RGB8 || op->out_size < 1) return PPMX_ERROR;
        if (op->dim == 0) nh = (uint32_t)op->out_size;  // ref:815-816
        else nw = (uint32_t)op->out_size;               // ref:841-842
        break;
    case PPMX_OP_CONV:
        if (layout != PPMX_LAYOUT_RGB8) return PPMX_ERROR;
        break;
    case PPMX_OP_LEVELS:
        if (layout == PPMX_LAYOUT_BITS) return PPMX_ERROR;
        break;
    case PPMX_OP_HIST_GRAY:
        if (layout != PPMX_LAYOUT_RGB8) return PPMX_ERROR;
        nw = nh = 0;
        break;
    default:
        return PPMX_ERROR;
    }
    if (ow) *ow = nw;
    if (oh) *oh = nh;
    if (olayout) *olayout = nl;
    return PPMX_OK;
}

static int upload_tables(const ppmx_op *op, DeviceTables *t, cudaStream_t s, bool async_pool)
{
    if (op->out_size < 1 || op->weights_sz < 1 || !op->weights || !op->indices) return fail("imresize: bad tables");
    size_t n = (size_t)op->out_size * op->weights_sz;
    size_t wbytes = n * sizeof(double), ibytes = n * sizeof(int);
    if (async_pool) CK(cudaMallocAsync(&t->base, wbytes + ibytes, s), "cudaMallocAsync tables");
    else CK(cudaMalloc(&t->base, wbytes + ibytes), "cudaMalloc tables");
    t->weights = (double *)t->base;
    t->indices = (int *)((uint8_t *)t->base + wbytes);
    CK(cudaMemcpyAsync(t->weights, op->weights, wbytes, cudaMemcpyHostToDevice, s), "upload weights");
    CK(cudaMemcpyAsync(t->indices, op->indices, ibytes, cudaMemcpyHostToDevice, s), "upload indices");
    return PPMX_OK;
}

// launches the kernel(s) of one operator on raw device pointers
static int launch_op(const ppmx_op *op, const uint8_t *d_src, uint32_t w, uint32_t h, int layout, uint8_t *d_dst,
                     const Band &band, unsigned long long *d_hist, const DeviceTables *tables, cudaStream_t s)
{
    const size_t npix = (size_t)w * h;
    const uint32_t y0 = band.full_h ? band.y0 : 0u;
    switch (op->kind) {
    case PPMX_OP_GRAY:
        CK(ppmx::gray(d_src, d_dst, npix, nullptr, s), "gray");
        return PPMX_OK;
    case PPMX_OP_GRAY_HIST:
        if (!d_hist) return fail("gray+hist: no histogram buffer");
        CK(ppmx::gray(d_src, d_dst, npix, d_hist, s), "gray+hist");
        return PPMX_OK;
    case PPMX_OP_HIST_GRAY:
        if (!d_hist) return fail("hist: no histogram buffer");
        CK(ppmx::hist_gray(d_src, npix, d_hist, s), "hist");
        return PPMX_OK;
    case PPMX_OP_MONO:
        CK(ppmx::mono_plane(d_src, d_dst, w, h, y0, s), "mono");
        return PPMX_OK;
    case PPMX_OP_MONO_BITS:
        CK(ppmx::mono_bits(d_src, d_dst, w, h, y0, s), "mono+pack");
        return PPMX_OK;
    case PPMX_OP_PACK_PBM:
        CK(ppmx::pack_pbm(d_src, layout == PPMX_LAYOUT_RGB8 ? 3 : 1, d_dst, w, h, s), "pack");
        return PPMX_OK;
    case PPMX_OP_EXTRACT_R:
        CK(ppmx::extract_r(d_src, d_dst, npix, s), "extract .r");
        return PPMX_OK;
    case PPMX_OP_FLIP:
        CK(ppmx::flip(d_src, d_dst, w, h, layout == PPMX_LAYOUT_RGB8 ? 3 : 1, op->flip_direction ? 1 : 0, s), "flip");
        return PPMX_OK;
    case PPMX_OP_ROTATE:
        if (op->angle_deg == 0) {
            CK(cudaMemcpyAsync(d_dst, d_src, npix * 3, cudaMemcpyDeviceToDevice, s), "rotate 0");
        } else if (op->angle_deg == 90 || op->angle_deg == 180 || op->angle_deg == 270) {
            CK(ppmx::rotate_orth(d_src, d_dst, w, h, op->angle_deg, s), "rotate");
        } else {
            CK(ppmx::rotate_bicubic(d_src, d_dst, w, h, op->new_width, op->new_height, op->cos_t, op->sin_t, s),
               "rotate (bicubic)");
        }
        return PPMX_OK;
    case PPMX_OP_IMRESIZE:
        if (!tables || !tables->weights) return fail("imresize: tables are not on the device");
        CK(ppmx::imresize(d_src, d_dst, w, h, op->out_size, op->dim, op->weights_sz, tables->weights, tables->indices, band, s),
           "imresize");
        return PPMX_OK;
    case PPMX_OP_CONV:
        if (!op->conv_coef) return fail("conv: no coefficients");
        CK(ppmx::conv(d_src, d_dst, w, h, op->conv_k, op->conv_coef, op->conv_div, op->conv_bias, band, s), "conv");
        return PPMX_OK;
    case PPMX_OP_LEVELS:
        if (!op->levels_lut) return fail("levels: no table");
        CK(ppmx::levels(d_src, d_dst, npix * (layout == PPMX_LAYOUT_RGB8 ? 3 : 1), op->levels_lut, s), "levels");
        return PPMX_OK;
    default:
        return fail("unknown operator kind");
    }
}

static int op_on(ppmx_gpu_ctx *c, int lane, const ppmx_op *op, const ppmx_gpu_image *src, ppmx_gpu_image **dst,
                 const DeviceTables *shared_tables, unsigned long long *d_hist)
{
    uint32_t ow, oh;
    int ol;
    if (ppmx_gpu_op_output(op, src->w, src->h, src->layout, &ow, &oh, &ol) != PPMX_OK)
        return fail("operator does not accept this raster layout");
    cudaStream_t s = c->lane[lane];
    ppmx_gpu_image *out = nullptr;
    if (op->kind != PPMX_OP_HIST_GRAY && image_alloc_on(c, lane, ow, oh, ol, &out) != PPMX_OK) return PPMX_ERROR;

    DeviceTables local;
    const DeviceTables *tables = shared_tables;
    if (op->kind == PPMX_OP_IMRESIZE && !tables) {
        if (upload_tables(op, &local, s, true) != PPMX_OK) {
            image_free_on(c, out);
            return PPMX_ERROR;
        }
        tables = &local;
    }
    if (d_hist) CK(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), s), "clear hist");
    // rotate's uncovered pixels are written as 0 by the kernel itself (ref:727); no memset needed
    int rc = launch_op(op, src->d, src->w, src->h, src->layout, out ? out->d : nullptr, Band(), d_hist, tables, s);
    if (local.base) cudaFreeAsync(local.base, s);
    if (rc != PPMX_OK) {
        image_free_on(c, out);
        return rc;
    }
    *dst = out;
    return PPMX_OK;
}

extern "C" int ppmx_gpu_op(ppmx_gpu_ctx *c, const ppmx_op *op, const ppmx_gpu_image *src, ppmx_gpu_image **dst,
                           uint64_t *hist_out)
{
    if (!c || !op || !src || !dst) return fail("ppmx_gpu_op: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    *dst = nullptr;
    const bool hist = (op->kind == PPMX_OP_HIST_GRAY || op->kind == PPMX_OP_GRAY_HIST);
    if (hist && !hist_out) return fail("histogram operator needs hist_out");
    int rc = op_on(c, src->lane, op, src, dst, nullptr, hist ? c->d_hist : nullptr);
    if (rc != PPMX_OK) return rc;
    if (hist) {
        cudaStream_t s = c->lane[src->lane];
        CK(cudaMemcpyAsync(c->h_hist, c->d_hist, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s), "hist D2H");
        CK(cudaStreamSynchronize(s), "sync");
        for (int i = 0; i < 256; i++) hist_out[i] = c->h_hist[i];
    }
    return PPMX_OK;
}

// converts (if needed) to the byte stream of `file_type` and starts the D2H copy; *tmp is a
// conversion raster the caller frees after the copy was enqueued
static int download_on(ppmx_gpu_ctx *c, int lane, const ppmx_gpu_image *img, int file_type, uint8_t *dst, size_t cap,
                       size_t *nbytes, ppmx_gpu_image **tmp)
{
    *tmp = nullptr;
    const ppmx_gpu_image *from = img;
    ppmx_op cv;
    memset(&cv, 0, sizeof(cv));
    if (file_type == PPMX_FILETYPE_PGM) {  // ref:263-267
        if (img->layout == PPMX_LAYOUT_RGB8) cv.kind = PPMX_OP_EXTRACT_R;
        else if (img->layout != PPMX_LAYOUT_R8) return fail("PGM output from a bit raster");
        else cv.kind = -1;
    } else if (file_type == PPMX_FILETYPE_PBM) {  // ref:268-284
        cv.kind = (img->layout == PPMX_LAYOUT_BITS) ? -1 : PPMX_OP_PACK_PBM;
    } else {  // ref:285-291
        if (img->layout != PPMX_LAYOUT_RGB8) return fail("PPM output needs an RGB raster");
        cv.kind = -1;
    }
    if (cv.kind >= 0) {
        if (op_on(c, lane, &cv, img, tmp, nullptr, nullptr) != PPMX_OK) return PPMX_ERROR;
        from = *tmp;
    }
    if (from->bytes > cap) return fail("destination buffer too small");
    if (from->bytes) CK(cudaMemcpyAsync(dst, from->d, from->bytes, cudaMemcpyDeviceToHost, c->lane[lane]), "download");
    if (nbytes) *nbytes = from->bytes;
    return PPMX_OK;
}

extern "C" int ppmx_gpu_download(ppmx_gpu_ctx *c, const ppmx_gpu_image *img, int file_type, uint8_t *dst, size_t cap,
                                 size_t *nbytes)
{
    if (!c || !img || (!dst && cap)) return fail("ppmx_gpu_download: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    ppmx_gpu_image *tmp = nullptr;
    int rc = download_on(c, img->lane, img, file_type, dst, cap, nbytes, &tmp);
    image_free_on(c, tmp);
    if (rc != PPMX_OK) return rc;
    CK(cudaStreamSynchronize(c->lane[img->lane]), "sync");
    return PPMX_OK;
}

// ---------------------------------------------------------------------------------------------
// the op chain (ref:1084-1155): buff / new_buff hand-over, then the writer's raster (ref:263-291)
// ---------------------------------------------------------------------------------------------

struct Chain {
    ppmx_gpu_ctx *c;
    int lane;
    ppmx_gpu_image *buff = nullptr, *newb = nullptr;  // newb may alias buff (flip, rotate 0)
    int file_type = PPMX_FILETYPE_PPM;
    int index = 0;  // position of this raster in a batch

    void drop_new()
    {  // the reference leaks a superseded new_buff; here it goes back to the pool
        if (newb && newb != buff) image_free_on(c, newb);
        newb = nullptr;
    }
    void renew()
    {  // renewBuffer, ref:1019-1026
        if (!newb) return;
        if (newb != buff) image_free_on(c, buff);
        buff = newb;
        newb = nullptr;
    }
    void release()
    {
        if (newb && newb != buff) image_free_on(c, newb);
        image_free_on(c, buff);
        buff = newb = nullptr;
    }
};

static int run_chain(Chain &ch, const ppmx_op *ops, int nops, const std::vector<DeviceTables> &tables)
{
    ppmx_gpu_ctx *c = ch.c;
    for (int i = 0; i < nops; i++) {
        ppmx_op op = ops[i];
        if (op.renew_before) ch.renew();
        ppmx_gpu_image *out = nullptr;
        const DeviceTables *t = (op.kind == PPMX_OP_IMRESIZE) ? &tables[i] : nullptr;
        switch (op.kind) {
        case PPMX_OP_MONO:
            // a bilevel result nothing else touches goes straight to packed bits (mono + ref:268-284)
            if (i == nops - 1) op.kind = PPMX_OP_MONO_BITS;
            /* fall through */
        case PPMX_OP_GRAY:
            if (op_on(c, ch.lane, &op, ch.buff, &out, nullptr, nullptr) != PPMX_OK) return PPMX_ERROR;
            ch.drop_new();
            ch.newb = out;
            ch.file_type = (op.kind == PPMX_OP_GRAY) ? PPMX_FILETYPE_PGM : PPMX_FILETYPE_PBM;  // ref:991, 956
            break;
        case PPMX_OP_GRAY_HIST: {  // extension: gray (ref:998-1000) + its 256-bin histogram in one pass
            if (!op.hist_out) return fail("gray+hist in a chain needs op.hist_out");
            unsigned long long *dh = nullptr;
            cudaStream_t s = c->lane[ch.lane];
            CK(cudaMallocAsync((void **)&dh, 256 * sizeof(unsigned long long), s), "cudaMallocAsync hist");
            int rc = op_on(c, ch.lane, &op, ch.buff, &out, nullptr, dh);
            if (rc == PPMX_OK)
                rc = cudaMemcpyAsync(op.hist_out + 256 * (size_t)ch.index, dh, 256 * sizeof(unsigned long long),
                                     cudaMemcpyDeviceToHost, s) == cudaSuccess ? PPMX_OK : fail("hist D2H");
            cudaFreeAsync(dh, s);
            if (rc != PPMX_OK) return rc;
            ch.drop_new();
            ch.newb = out;
            ch.file_type = PPMX_FILETYPE_PGM;
            break;
        }
        case PPMX_OP_FLIP:
            // ref:896: works on buff itself and aliases new_buff to it
            if (op_on(c, ch.lane, &op, ch.buff, &out, nullptr, nullptr) != PPMX_OK) return PPMX_ERROR;
            ch.drop_new();
            image_free_on(c, ch.buff);
            ch.buff = ch.newb = out;
            break;
        case PPMX_OP_ROTATE:
            if (op.angle_deg == 0) {  // ref:701-705
                ch.drop_new();
                ch.newb = ch.buff;
                break;
            }
            /* fall through */
        case PPMX_OP_IMRESIZE:
        case PPMX_OP_CONV:
        case PPMX_OP_LEVELS:
            if (op_on(c, ch.lane, &op, ch.buff, &out, t, nullptr) != PPMX_OK) return PPMX_ERROR;
            ch.drop_new();
            ch.newb = out;
            break;
        default:
            return fail("operator not allowed in a chain");
        }
    }
    if (!ch.newb) return fail("Error: no data to write");  // ref:235
    return PPMX_OK;
}

static int prepare_tables(ppmx_gpu_ctx *c, const ppmx_op *ops, int nops, std::vector<DeviceTables> &tables)
{
    tables.assign(nops, DeviceTables());
    bool any = false;
    for (int i = 0; i < nops; i++)
        if (ops[i].kind == PPMX_OP_IMRESIZE) {
            if (upload_tables(&ops[i], &tables[i], c->lane[0], true) != PPMX_OK) return PPMX_ERROR;
            any = true;
        }
    if (any) {
        CK(cudaEventRecord(c->tables_ready, c->lane[0]), "event record");
        for (int l = 1; l < kLanes; l++) CK(cudaStreamWaitEvent(c->lane[l], c->tables_ready, 0), "event wait");
    }
    return PPMX_OK;
}

static void release_tables(ppmx_gpu_ctx *c, std::vector<DeviceTables> &tables)
{
    // every lane has been synchronised by the caller
    for (auto &t : tables)
        if (t.base) cudaFreeAsync(t.base, c->lane[0]);
    tables.clear();
}

extern "C" int ppmx_gpu_apply_batch(ppmx_gpu_ctx *c, const ppmx_op *ops, int nops, const uint8_t *src, uint32_t w,
                                    uint32_t h, int count, uint8_t *dst, size_t dst_stride, size_t *dst_bytes_each,
                                    uint32_t *out_w, uint32_t *out_h, int *out_file_type)
{
    if (!c || !ops || nops < 1 || !src || !dst || count < 1) return fail("ppmx_gpu_apply: bad argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    std::vector<DeviceTables> tables;
    if (prepare_tables(c, ops, nops, tables) != PPMX_OK) return PPMX_ERROR;

    const size_t in_bytes = (size_t)w * h * 3;
    int rc = PPMX_OK;
    size_t each = 0;
    uint32_t ow = 0, oh = 0;
    int ft = PPMX_FILETYPE_PPM;
    // at most kInFlight rasters per lane are enqueued ahead of the GPU, so the stream-ordered pool
    // holds a bounded number of rasters however long the batch is
    constexpr int kInFlight = 4;
    cudaEvent_t done[kLanes][kInFlight] = {};
    for (int i = 0; i < count && rc == PPMX_OK; i++) {
        Chain ch;
        ch.c = c;
        ch.lane = i % kLanes;
        ch.index = i;
        const int slot = (i / kLanes) % kInFlight;
        if (done[ch.lane][slot]) CK(cudaEventSynchronize(done[ch.lane][slot]), "event sync");
        else CK(cudaEventCreateWithFlags(&done[ch.lane][slot], cudaEventDisableTiming), "event create");
        rc = upload_on(c, ch.lane, src + (size_t)i * in_bytes, w, h, PPMX_LAYOUT_RGB8, &ch.buff);
        if (rc == PPMX_OK) rc = run_chain(ch, ops, nops, tables);
        if (rc == PPMX_OK) {
            ppmx_gpu_image *tmp = nullptr;
            rc = download_on(c, ch.lane, ch.newb, ch.file_type, dst + (size_t)i * dst_stride, dst_stride, &each, &tmp);
            image_free_on(c, tmp);
            ow = ch.newb->w;
            oh = ch.newb->h;
            ft = ch.file_type;
        }
        ch.release();
        cudaEventRecord(done[ch.lane][slot], c->lane[ch.lane]);
    }
    for (int l = 0; l < kLanes; l++) {
        cudaError_t e = cudaStreamSynchronize(c->lane[l]);
        if (e != cudaSuccess && rc == PPMX_OK) rc = fail("stream sync", e);
        for (int k = 0; k < kInFlight; k++)
            if (done[l][k]) cudaEventDestroy(done[l][k]);
    }
    release_tables(c, tables);
    if (rc != PPMX_OK) return rc;
    if (dst_bytes_each) *dst_bytes_each = each;
    if (out_w) *out_w = ow;
    if (out_h) *out_h = oh;
    if (out_file_type) *out_file_type = ft;
    return PPMX_OK;
}

extern "C" int ppmx_gpu_apply(ppmx_gpu_ctx *c, const ppmx_op *ops, int nops, const uint8_t *src, uint32_t w, uint32_t h,
                              uint8_t *dst, size_t dst_cap, size_t *dst_bytes, uint32_t *out_w, uint32_t *out_h,
                              int *out_file_type)
{
    return ppmx_gpu_apply_batch(c, ops, nops, src, w, h, 1, dst, dst_cap, dst_bytes, out_w, out_h, out_file_type);
}

// ---------------------------------------------------------------------------------------------
// raw launches on caller-owned device memory
// ---------------------------------------------------------------------------------------------

extern "C" int ppmx_gpu_tables_upload(const ppmx_op *op, void **d_tables)
{
    if (!op || !d_tables) return fail("ppmx_gpu_tables_upload: null argument");
    DeviceTables t;
    if (upload_tables(op, &t, 0, false) != PPMX_OK) return PPMX_ERROR;
    CK(cudaStreamSynchronize(0), "sync");
    *d_tables = t.base;
    return PPMX_OK;
}

extern "C" void ppmx_gpu_tables_free(void *d_tables)
{
    if (d_tables) cudaFree(d_tables);
}

extern "C" int ppmx_gpu_launch(const ppmx_op *op, const void *d_src, uint32_t w, uint32_t h, int src_layout, void *d_dst,
                               const ppmx_band *band, void *d_hist, void *d_tables, void *stream)
{
    if (!op || !d_src) return fail("ppmx_gpu_launch: null argument");
    Band b;
    if (band && band->full_h) {
        b.full_h = band->full_h;
        b.y0 = band->y0;
        b.top = (const uint8_t *)band->d_top;
        b.bottom = (const uint8_t *)band->d_bottom;
        b.halo = band->halo;
        b.out_y0 = band->out_y0;
        b.out_rows = band->out_rows;
    }
    DeviceTables t;
    if (op->kind == PPMX_OP_IMRESIZE) {
        if (!d_tables) return fail("imresize: d_tables missing");
        size_t n = (size_t)op->out_size * op->weights_sz;
        t.base = d_tables;
        t.weights = (double *)d_tables;
        t.indices = (int *)((uint8_t *)d_tables + n * sizeof(double));
    }
    return launch_op(op, (const uint8_t *)d_src, w, h, src_layout, (uint8_t *)d_dst, b, (unsigned long long *)d_hist, &t,
                     (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: CUDA IPC so that one process per GPU can read a neighbour's band over NVLink
// ---------------------------------------------------------------------------------------------

extern "C" int ppmx_gpu_device_alloc(ppmx_gpu_ctx *c, size_t bytes, void **device_ptr)
{
    if (!c || !device_ptr) return fail("ppmx_gpu_device_alloc: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    CK(cudaMalloc(device_ptr, bytes ? bytes : 16), "cudaMalloc");  // plain cudaMalloc: IPC-exportable
    return PPMX_OK;
}

extern "C" void ppmx_gpu_device_free(ppmx_gpu_ctx *c, void *device_ptr)
{
    if (c) cudaSetDevice(c->device);
    if (device_ptr) cudaFree(device_ptr);
}

extern "C" int ppmx_gpu_copy(ppmx_gpu_ctx *c, void *dst, const void *src, size_t bytes, int kind)
{
    if (!c || (bytes && (!dst || !src))) return fail("ppmx_gpu_copy: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, k, c->lane[0]), "cudaMemcpyAsync");
    CK(cudaStreamSynchronize(c->lane[0]), "sync");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_ipc_export(ppmx_gpu_ctx *c, const void *device_ptr, uint8_t handle[64])
{
    if (!c || !device_ptr || !handle) return fail("ppmx_gpu_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, const_cast<void *>(device_ptr)), "cudaIpcGetMemHandle");
    memcpy(handle, &hd, 64);
    return PPMX_OK;
}

extern "C" int ppmx_gpu_ipc_open(ppmx_gpu_ctx *c, const uint8_t handle[64], void **device_ptr)
{
    if (!c || !device_ptr || !handle) return fail("ppmx_gpu_ipc_open: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    CK(cudaIpcOpenMemHandle(device_ptr, hd, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_ipc_close(ppmx_gpu_ctx *c, void *device_ptr)
{
    if (!c || !device_ptr) return fail("ppmx_gpu_ipc_close: null argument");
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    CK(cudaIpcCloseMemHandle(device_ptr), "cudaIpcCloseMemHandle");
    return PPMX_OK;
}
