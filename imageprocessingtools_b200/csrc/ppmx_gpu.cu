// ppmx_gpu.cu -- the C ABI of include/ppmx_gpu.h: contexts, rasters in HBM, single operators, raw launches, CUDA
// graphs and IPC.  The op chain (ref:1084-1155 = /root/reference/ppmx-edward.c), its row parts and the multi-device
// split live in ppmx_chain.cu; all arithmetic lives in ppmx_{color,geometry,bicubic,conv}.cu.  There is no CPU
// fallback anywhere in this library.
#include "ppmx_ctx.h"

#include <cstdlib>
#include <cstring>
#include <new>

using ppmx::Band;
using ppmx::DeviceTables;
using ppmx::fail;
using ppmx::image_alloc_on;
using ppmx::image_free_on;
using ppmx::kLanes;
using ppmx::launch_op;
using ppmx::primary;
using ppmx::upload_tables;

#define PPMX_VERSION "ppmx-b200 0.2 (sm_100a)"
#define CK PPMX_CK

int ppmx::fail(const char *what, cudaError_t e)
{
    if (e != cudaSuccess) printf("ppmx_gpu error: %s: %s\n", what, cudaGetErrorString(e));
    else printf("ppmx_gpu error: %s\n", what);
    fflush(stdout);
    return PPMX_ERROR;
}

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
    if (key && !strcmp(key, "variant")) {
#ifdef PPMX_TUNING
        ppmx::g_variant = value;
#else
        // the release library carries the default (best measured) kernels only; the alternative implementations
        // live in libppmx_gpu_tuning.so (same sources, -DPPMX_TUNING), which tools/sweep.py and the variant tests load
        if (value != 0) return PPMX_ERROR;
#endif
    } else if (key && !strcmp(key, "pdl")) ppmx::g_pdl.store(value ? 1 : 0);
    else return PPMX_ERROR;
    return PPMX_OK;
}

extern "C" const char *ppmx_gpu_version(void)
{
#ifdef PPMX_TUNING
    return PPMX_VERSION " tuning";
#else
    return PPMX_VERSION;
#endif
}
extern "C" uint64_t ppmx_gpu_launch_count(void) { return ppmx::launch_count(); }

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------

static int init_one(ppmx_gpu_ctx *c, int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail("no CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= n) return fail("ppmx_gpu_init: no such device");
    CK(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10) return fail("this build carries sm_100a code only and needs a B200-class device");
    c->device = device;
    for (int i = 0; i < kLanes; i++) {
        CK(cudaStreamCreateWithFlags(&c->lane[i], cudaStreamNonBlocking), "cudaStreamCreate");
        CK(cudaEventCreateWithFlags(&c->lane_done[i], cudaEventDisableTiming), "cudaEventCreate");
    }
    CK(cudaEventCreateWithFlags(&c->tables_ready, cudaEventDisableTiming), "cudaEventCreate");
    CK(cudaMalloc(&c->d_hist, 256 * sizeof(unsigned long long)), "cudaMalloc hist");
    CK(cudaHostAlloc(&c->h_hist, 256 * sizeof(unsigned long long), cudaHostAllocPortable), "cudaHostAlloc hist");
    // a pool of the context's own (freed rasters stay cached in it; the device's default pool is left alone)
    cudaMemPoolProps pp;
    memset(&pp, 0, sizeof(pp));
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    CK(cudaMemPoolCreate(&c->pool, &pp), "cudaMemPoolCreate");
    unsigned long long keep = ~0ull;
    CK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep), "cudaMemPoolSetAttribute");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_init(ppmx_gpu_ctx **out, int device)
{
    if (!out) return fail("ppmx_gpu_init: null ctx pointer");
    *out = nullptr;
    ppmx_gpu_ctx *c = new (std::nothrow) ppmx_gpu_ctx();
    if (!c) return fail("out of host memory");
    if (init_one(c, device) != PPMX_OK) {
        ppmx_gpu_free(c);  // whatever was created so far
        return PPMX_ERROR;
    }
    *out = c;
    return PPMX_OK;
}

extern "C" int ppmx_gpu_init_multi(ppmx_gpu_ctx **out, const int *devices, int ndev)
{
    if (!out) return fail("ppmx_gpu_init_multi: null ctx pointer");
    *out = nullptr;
    if (ndev < 0 || ndev > 64 || (ndev > 0 && !devices)) return fail("ppmx_gpu_init_multi: bad device list");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail("no CUDA device (this library has no CPU fallback)", e);
    if (ndev == 0) ndev = n;  // every visible device
    ppmx_gpu_ctx *m = new (std::nothrow) ppmx_gpu_ctx();
    if (!m) return fail("out of host memory");
    for (int i = 0; i < ndev; i++) {
        ppmx_gpu_ctx *c = nullptr;
        if (ppmx_gpu_init(&c, devices ? devices[i] : i) != PPMX_OK) {
            ppmx_gpu_free(m);
            return PPMX_ERROR;
        }
        m->children.push_back(c);
    }
    m->device = m->children[0]->device;
    *out = m;
    return PPMX_OK;
}

extern "C" int ppmx_gpu_device_count(const ppmx_gpu_ctx *c)
{
    if (!c) return 0;
    return c->children.empty() ? 1 : (int)c->children.size();
}

extern "C" void ppmx_gpu_free(ppmx_gpu_ctx *c)
{
    if (!c) return;
    for (ppmx_gpu_ctx *k : c->children) ppmx_gpu_free(k);
    if (c->children.empty()) {
        cudaSetDevice(c->device);
        for (int i = 0; i < kLanes; i++) {
            if (c->lane[i]) {
                cudaStreamSynchronize(c->lane[i]);
                cudaStreamDestroy(c->lane[i]);
            }
            if (c->lane_done[i]) cudaEventDestroy(c->lane_done[i]);
        }
        if (c->tables_ready) cudaEventDestroy(c->tables_ready);
        if (c->pool) cudaMemPoolDestroy(c->pool);  // hands the cached rasters back to the driver
        if (c->d_hist) cudaFree(c->d_hist);
        if (c->h_hist) cudaFreeHost(c->h_hist);
    }
    delete c;
}

extern "C" void *ppmx_gpu_host_alloc(ppmx_gpu_ctx *c, size_t bytes)
{
    void *p = nullptr;
    if (c) cudaSetDevice(primary(c)->device);
    // portable: a multi-device context copies to and from the same host raster on every device
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        fail("cudaHostAlloc");
        return nullptr;
    }
    return p;
}

extern "C" void ppmx_gpu_host_free(ppmx_gpu_ctx *c, void *p)
{
    (void)c;
    if (p) cudaFreeHost(p);
}

extern "C" int ppmx_gpu_host_register(ppmx_gpu_ctx *c, void *p, size_t bytes)
{
    if (!p || !bytes) return fail("ppmx_gpu_host_register: null argument");
    if (c) cudaSetDevice(primary(c)->device);
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable), "cudaHostRegister");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_host_unregister(ppmx_gpu_ctx *c, void *p)
{
    (void)c;
    if (!p) return fail("ppmx_gpu_host_unregister: null argument");
    CK(cudaHostUnregister(p), "cudaHostUnregister");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_sync(ppmx_gpu_ctx *c)
{
    if (!c) return fail("null ctx");
    if (!c->children.empty()) {
        for (ppmx_gpu_ctx *k : c->children)
            if (ppmx_gpu_sync(k) != PPMX_OK) return PPMX_ERROR;
        return PPMX_OK;
    }
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    for (int i = 0; i < kLanes; i++) CK(cudaStreamSynchronize(c->lane[i]), "cudaStreamSynchronize");
    return PPMX_OK;
}

// ---------------------------------------------------------------------------------------------
// rasters in HBM
// ---------------------------------------------------------------------------------------------

cudaError_t ppmx::pool_alloc(ppmx_gpu_ctx *c, void **p, size_t bytes, cudaStream_t s)
{
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 16, c->pool, s);
}

int ppmx::image_alloc_on(ppmx_gpu_ctx *c, int lane, uint32_t w, uint32_t h, int layout, ppmx_gpu_image **out)
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
    cudaError_t e = pool_alloc(c, (void **)&im->d, bytes + 16, c->lane[lane]);
    if (e != cudaSuccess) {
        delete im;
        return fail("can not allocate image buff in HBM", e);  // wording of ref:926
    }
    *out = im;
    return PPMX_OK;
}

void ppmx::image_free_on(ppmx_gpu_ctx *c, ppmx_gpu_image *im)
{
    if (!im) return;
    if (im->d) cudaFreeAsync(im->d, c->lane[im->lane]);
    delete im;
}

extern "C" int ppmx_gpu_image_alloc(ppmx_gpu_ctx *c, uint32_t w, uint32_t h, int layout, ppmx_gpu_image **img)
{
    if (!c || !img) return fail("ppmx_gpu_image_alloc: null argument");
    c = primary(c);
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    return image_alloc_on(c, 0, w, h, layout, img);
}

extern "C" void ppmx_gpu_image_free(ppmx_gpu_ctx *c, ppmx_gpu_image *img)
{
    if (!c) return;
    c = primary(c);
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

extern "C" int ppmx_gpu_upload(ppmx_gpu_ctx *c, const uint8_t *src, uint32_t w, uint32_t h, int layout,
                               ppmx_gpu_image **img)
{
    if (!c || !img || (!src && w && h)) return fail("ppmx_gpu_upload: null argument");
    c = primary(c);
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    *img = nullptr;
    ppmx_gpu_image *im = nullptr;
    if (image_alloc_on(c, 0, w, h, layout, &im) != PPMX_OK) return PPMX_ERROR;
    if (im->bytes) {
        cudaError_t e = cudaMemcpyAsync(im->d, src, im->bytes, cudaMemcpyHostToDevice, c->lane[0]);
        if (e != cudaSuccess) {
            image_free_on(c, im);
            return fail("upload", e);
        }
    }
    *img = im;
    return PPMX_OK;
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

int ppmx::upload_tables(ppmx_gpu_ctx *c, const ppmx_op *op, DeviceTables *t, cudaStream_t s)
{
    if (op->out_size < 1 || op->weights_sz < 1 || !op->weights || !op->indices) return fail("imresize: bad tables");
    size_t n = (size_t)op->out_size * op->weights_sz;
    size_t wbytes = n * sizeof(double), ibytes = n * sizeof(int);
    if (c) CK(pool_alloc(c, &t->base, wbytes + ibytes, s), "pool alloc (tables)");
    else CK(cudaMalloc(&t->base, wbytes + ibytes), "cudaMalloc tables");
    t->weights = (double *)t->base;
    t->indices = (int *)((uint8_t *)t->base + wbytes);
    cudaError_t e = cudaMemcpyAsync(t->weights, op->weights, wbytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->indices, op->indices, ibytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
        if (c) cudaFreeAsync(t->base, s);
        else cudaFree(t->base);
        *t = DeviceTables();
        return fail("upload tables", e);
    }
    return PPMX_OK;
}

// launches the kernel(s) of one operator on raw device pointers
int ppmx::launch_op(const ppmx_op *op, const uint8_t *d_src, uint32_t w, uint32_t h, int layout, uint8_t *d_dst,
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
                 unsigned long long *d_hist)
{
    uint32_t ow, oh;
    int ol;
    if (ppmx_gpu_op_output(op, src->w, src->h, src->layout, &ow, &oh, &ol) != PPMX_OK)
        return fail("operator does not accept this raster layout");
    cudaStream_t s = c->lane[lane];
    ppmx_gpu_image *out = nullptr;
    if (op->kind != PPMX_OP_HIST_GRAY && image_alloc_on(c, lane, ow, oh, ol, &out) != PPMX_OK) return PPMX_ERROR;

    DeviceTables local;
    int rc = PPMX_OK;
    if (op->kind == PPMX_OP_IMRESIZE) rc = upload_tables(c, op, &local, s);
    if (rc == PPMX_OK && d_hist && cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), s) != cudaSuccess)
        rc = fail("clear hist");
    // rotate's uncovered pixels are written as 0 by the kernel itself (ref:727); no memset needed
    if (rc == PPMX_OK)
        rc = launch_op(op, src->d, src->w, src->h, src->layout, out ? out->d : nullptr, Band(), d_hist, &local, s);
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
    c = primary(c);
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    *dst = nullptr;
    const bool hist = (op->kind == PPMX_OP_HIST_GRAY || op->kind == PPMX_OP_GRAY_HIST);
    if (hist && !hist_out) return fail("histogram operator needs hist_out");
    int rc = op_on(c, src->lane, op, src, dst, hist ? c->d_hist : nullptr);
    if (rc != PPMX_OK) return rc;
    if (hist) {
        cudaStream_t s = c->lane[src->lane];
        cudaError_t e = cudaMemcpyAsync(c->h_hist, c->d_hist, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            image_free_on(c, *dst);
            *dst = nullptr;
            return fail("hist D2H", e);
        }
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
        if (op_on(c, lane, &cv, img, tmp, nullptr) != PPMX_OK) return PPMX_ERROR;
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
    c = primary(c);
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    ppmx_gpu_image *tmp = nullptr;
    int rc = download_on(c, img->lane, img, file_type, dst, cap, nbytes, &tmp);
    image_free_on(c, tmp);
    cudaError_t e = cudaStreamSynchronize(c->lane[img->lane]);  // also after a failure: nothing may still write to dst
    if (rc != PPMX_OK) return rc;
    if (e != cudaSuccess) return fail("sync", e);
    return PPMX_OK;
}

// ---------------------------------------------------------------------------------------------
// raw launches on caller-owned device memory
// ---------------------------------------------------------------------------------------------

extern "C" int ppmx_gpu_tables_upload(const ppmx_op *op, void **d_tables)
{
    if (!op || !d_tables) return fail("ppmx_gpu_tables_upload: null argument");
    DeviceTables t;
    if (upload_tables(nullptr, op, &t, 0) != PPMX_OK) return PPMX_ERROR;
    cudaError_t e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        cudaFree(t.base);
        return fail("sync", e);
    }
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
// CUDA graphs: a sequence of ppmx_gpu_launch calls (they only enqueue kernels) recorded once and replayed with ONE
// driver call per step, so the host's launch cost is off the device's critical path
// ---------------------------------------------------------------------------------------------

struct ppmx_gpu_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    unsigned long long kernels = 0;
};

extern "C" int ppmx_gpu_graph_begin(void *stream)
{
    CK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_graph_end(void *stream, ppmx_gpu_graph **out, uint64_t *kernel_nodes)
{
    if (!out) return fail("ppmx_gpu_graph_end: null argument");
    *out = nullptr;
    ppmx_gpu_graph *g = new (std::nothrow) ppmx_gpu_graph();
    if (!g) return fail("out of host memory");
    cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &g->graph);
    if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e == cudaSuccess) {
        size_t n = 0;
        e = cudaGraphGetNodes(g->graph, nullptr, &n);
        if (e == cudaSuccess && n) {
            std::vector<cudaGraphNode_t> nodes(n);
            e = cudaGraphGetNodes(g->graph, nodes.data(), &n);
            for (size_t i = 0; e == cudaSuccess && i < n; i++) {
                cudaGraphNodeType t;
                e = cudaGraphNodeGetType(nodes[i], &t);
                if (e == cudaSuccess && t == cudaGraphNodeTypeKernel) g->kernels++;
            }
        }
    }
    if (e != cudaSuccess) {
        ppmx_gpu_graph_free(g);
        return fail("graph capture", e);
    }
    if (kernel_nodes) *kernel_nodes = g->kernels;
    *out = g;
    return PPMX_OK;
}

extern "C" int ppmx_gpu_graph_launch(ppmx_gpu_graph *g, void *stream)
{
    if (!g || !g->exec) return fail("ppmx_gpu_graph_launch: null graph");
    CK(cudaGraphLaunch(g->exec, (cudaStream_t)stream), "cudaGraphLaunch");
    ppmx::add_launches(g->kernels);
    return PPMX_OK;
}

extern "C" void ppmx_gpu_graph_free(ppmx_gpu_graph *g)
{
    if (!g) return;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: CUDA IPC so that one process per GPU can read a neighbour's band over NVLink
// ---------------------------------------------------------------------------------------------

extern "C" int ppmx_gpu_device_alloc(ppmx_gpu_ctx *c, size_t bytes, void **device_ptr)
{
    if (!c || !device_ptr) return fail("ppmx_gpu_device_alloc: null argument");
    CK(cudaSetDevice(primary(c)->device), "cudaSetDevice");
    CK(cudaMalloc(device_ptr, bytes ? bytes : 16), "cudaMalloc");  // plain cudaMalloc: IPC-exportable
    return PPMX_OK;
}

extern "C" void ppmx_gpu_device_free(ppmx_gpu_ctx *c, void *device_ptr)
{
    if (c) cudaSetDevice(primary(c)->device);
    if (device_ptr) cudaFree(device_ptr);
}

extern "C" int ppmx_gpu_copy(ppmx_gpu_ctx *c, void *dst, const void *src, size_t bytes, int kind)
{
    if (!c || (bytes && (!dst || !src))) return fail("ppmx_gpu_copy: null argument");
    c = primary(c);
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
    CK(cudaSetDevice(primary(c)->device), "cudaSetDevice");
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, const_cast<void *>(device_ptr)), "cudaIpcGetMemHandle");
    memcpy(handle, &hd, 64);
    return PPMX_OK;
}

extern "C" int ppmx_gpu_ipc_open(ppmx_gpu_ctx *c, const uint8_t handle[64], void **device_ptr)
{
    if (!c || !device_ptr || !handle) return fail("ppmx_gpu_ipc_open: null argument");
    CK(cudaSetDevice(primary(c)->device), "cudaSetDevice");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    CK(cudaIpcOpenMemHandle(device_ptr, hd, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return PPMX_OK;
}

extern "C" int ppmx_gpu_ipc_close(ppmx_gpu_ctx *c, void *device_ptr)
{
    if (!c || !device_ptr) return fail("ppmx_gpu_ipc_close: null argument");
    CK(cudaSetDevice(primary(c)->device), "cudaSetDevice");
    CK(cudaIpcCloseMemHandle(device_ptr), "cudaIpcCloseMemHandle");
    return PPMX_OK;
}
