// ppmx_kernels.cu -- hand-written sm_100a kernels for the per-pixel loops of ppmx-edward.c.
//
// "ref:N" = /root/reference/ppmx-edward.c line N.  Rasters are flat, packed, row-major:
// RGB8 = 3 bytes per pixel (the reference's `pixel`, ref:39-43), R8 = the .r member only.
// Everything integer is done in integers; the two bicubic operators run in FP64 with
// explicit round-to-nearest multiplies and adds (never an FMA) in the reference's order.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
#include "ppmx_kernels.h"

#include <cuda.h>

namespace ppmx {

int g_variant = 0;
static unsigned long long g_launches = 0;
unsigned long long launch_count() { return g_launches; }

static int g_sm_count = 0;
static int sm_count()
{
    if (!g_sm_count) {  // every device a process sees in this pool is the same B200 part
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sm_count <= 0)
            g_sm_count = 148;
    }
    return g_sm_count;
}

#define PPMX_LAUNCHED() (cudaGetLastError())

// Programmatic dependent launch (sm_90+): every kernel in this file starts with pdl_wait(), so a
// launch may be scheduled while its predecessor in the stream is still draining; only index
// arithmetic runs before the wait, all global-memory traffic after it.
int g_pdl = 1;  // ppmx_gpu_set_tuning("pdl", 0/1)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define PDL_PROLOGUE() \
    do {               \
        pdl_trigger(); \
        pdl_wait();    \
    } while (0)

// opt in to > 48 KB dynamic shared memory once per (kernel, device)
template <typename K>
static void allow_smem(K kernel, size_t bytes, bool (&done)[64])
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev]) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    done[dev] = true;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ++g_launches;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned4(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// grid size for a grid-stride kernel: whole waves of the 148 SMs, capped by the work
static inline unsigned wave_grid(size_t work_items, unsigned block, unsigned ctas_per_sm)
{
    size_t need = (work_items + block - 1) / block;
    size_t wave = (size_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < wave ? need : wave);
}

// ------------------------------------------------------------------------------------------
// integer helpers
// ------------------------------------------------------------------------------------------

// s / 3 for 0 <= s <= 765, exact (checked exhaustively in tests/test_host_logic.py)
__device__ __forceinline__ uint32_t div3(uint32_t s) { return (s * 43691u) >> 17; }
// (one multiply-high, __umulhi(s, 0x55555556), is also exact but measured 10% slower: IMAD.HI
// is not a full-rate instruction)

// greys of 4 consecutive pixels held in 3 little-endian words (12 bytes r0 g0 b0 r1 ...), ref:1000
__device__ __forceinline__ void gray4_split(uint32_t a, uint32_t b, uint32_t c, uint32_t (&q)[4])
{
    q[0] = div3(__dp4a(a, 0x00010101u, 0u));
    q[1] = div3(__dp4a(a, 0x01000000u, __dp4a(b, 0x00000101u, 0u)));
    q[2] = div3(__dp4a(b, 0x01010000u, __dp4a(c, 0x00000001u, 0u)));
    q[3] = div3(__dp4a(c, 0x01010100u, 0u));
}
__device__ __forceinline__ uint32_t pack4(const uint32_t (&q)[4])
{  // one byte per pixel, pixel 0 in the low byte
    return __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
}
__device__ __forceinline__ uint32_t gray4(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t q[4];
    gray4_split(a, b, c, q);
    return pack4(q);
}

// 16 pixels = 48 bytes = three 16-byte vectors -> 16 grey bytes
__device__ __forceinline__ uint4 gray16(uint4 p, uint4 q, uint4 r)
{
    uint4 o;
    o.x = gray4(p.x, p.y, p.z);
    o.y = gray4(p.w, q.x, q.y);
    o.z = gray4(q.z, q.w, r.x);
    o.w = gray4(r.y, r.z, r.w);
    return o;
}

// Bayer thresholds of ref:954 times 255 (all exact), in the reference's own index order
// (x%4)*4 + (y%4), ref:967.  bit = grey < threshold  <=>  !(grey >= matrix*255).
__constant__ uint8_t c_bayer[16] = {32, 255, 48, 208, 160, 96, 176, 112, 64, 224, 16, 240, 192, 128, 144, 80};

// thresholds for 4 consecutive pixels starting at x % 4 == 0 on row y, packed like gray4's result
__device__ __forceinline__ uint32_t bayer_row4(uint32_t y)
{
    uint32_t yy = y & 3u;
    return (uint32_t)c_bayer[yy] | ((uint32_t)c_bayer[4 + yy] << 8) | ((uint32_t)c_bayer[8 + yy] << 16) |
           ((uint32_t)c_bayer[12 + yy] << 24);
}

// 4 packed greys vs 4 packed thresholds -> nibble, pixel 0 in bit 3 (MSB first, ref:273)
__device__ __forceinline__ uint32_t mono_nibble(uint32_t g4, uint32_t t4)
{
    uint32_t m = __vcmpltu4(g4, t4) & 0x01010101u;  // byte i = 1 iff grey_i < thr_i
    return ((m * 0x08040201u) >> 24) & 0xFu;        // b0<<3 | b1<<2 | b2<<1 | b3
}

// ------------------------------------------------------------------------------------------
// gray  (ref:998-1000)  RGB8 -> R8, flat over the raster; optional fused histogram (extension)
// ------------------------------------------------------------------------------------------

// per-CTA histogram: one 256-bin copy per warp in shared memory, merged once at the end
template <int WARPS>
struct SmemHist {
    uint32_t bins[WARPS][256];
    __device__ void clear()
    {
        for (int i = threadIdx.x; i < WARPS * 256; i += blockDim.x) (&bins[0][0])[i] = 0;
    }
    __device__ __forceinline__ void add4(uint32_t g4)
    {
        uint32_t *b = bins[threadIdx.x >> 5];
        atomicAdd(&b[g4 & 0xFF], 1u);
        atomicAdd(&b[(g4 >> 8) & 0xFF], 1u);
        atomicAdd(&b[(g4 >> 16) & 0xFF], 1u);
        atomicAdd(&b[g4 >> 24], 1u);
    }
    __device__ __forceinline__ void add1(uint32_t g) { atomicAdd(&bins[threadIdx.x >> 5][g & 0xFF], 1u); }
    __device__ void flush(unsigned long long *d_hist)
    {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            unsigned long long t = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) t += bins[w][i];
            if (t) atomicAdd(&d_hist[i], t);
        }
    }
};

template <bool HIST, bool STORE>
__global__ void __launch_bounds__(256) gray_vec_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                       size_t ngroups, size_t npix, unsigned long long *d_hist)
{
    PDL_PROLOGUE();
    __shared__ SmemHist<HIST ? 8 : 1> sh;
    if (HIST) {
        sh.clear();
        __syncthreads();
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        const uint4 *p = src + 3 * g;
        uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        uint4 o = gray16(a, b, c);
        if (STORE) dst[g] = o;
        if (HIST) {
            sh.add4(o.x);
            sh.add4(o.y);
            sh.add4(o.z);
            sh.add4(o.w);
        }
    }
    // the last (npix % 16) pixels, scalar
    if (blockIdx.x == 0 && threadIdx.x < (npix - ngroups * 16)) {
        size_t i = ngroups * 16 + threadIdx.x;
        const uint8_t *s8 = reinterpret_cast<const uint8_t *>(src) + 3 * i;
        uint32_t g = div3((uint32_t)s8[0] + s8[1] + s8[2]);
        if (STORE) reinterpret_cast<uint8_t *>(dst)[i] = (uint8_t)g;
        if (HIST) sh.add1(g);
    }
    if (HIST) {
        __syncthreads();
        sh.flush(d_hist);
    }
}

// any alignment: one pixel per thread
template <bool HIST, bool STORE>
__global__ void __launch_bounds__(256) gray_scalar_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                          size_t npix, unsigned long long *d_hist)
{
    PDL_PROLOGUE();
    __shared__ SmemHist<HIST ? 8 : 1> sh;
    if (HIST) {
        sh.clear();
        __syncthreads();
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        uint32_t g = div3((uint32_t)src[3 * i] + src[3 * i + 1] + src[3 * i + 2]);
        if (STORE) dst[i] = (uint8_t)g;
        if (HIST) sh.add1(g);
    }
    if (HIST) {
        __syncthreads();
        sh.flush(d_hist);
    }
}

// ---- histogram with contention-free, thread-private byte counters ---------------------------
// Each thread owns one 4-byte column in each of 64 shared-memory rows: byte (bin & 3) of row
// (bin >> 2).  A lane therefore always hits its own bank, needs no atomics and takes the same
// time on a constant image as on noise.  A byte counter holds 255, so after at most 15 groups
// of 16 pixels per thread the CTA folds the counters into 256 per-CTA totals (dp4a column sums)
// and clears them; the totals go to global memory in one atomic pass at the end.
constexpr int HP_THREADS = 256;
constexpr int HP_MAX_GROUPS = 15;
constexpr size_t HP_SMEM = (64 * HP_THREADS + 256) * sizeof(uint32_t);

__device__ __forceinline__ void hp_bump(uint8_t *mine, uint32_t g)
{
    uint8_t *p = mine + ((g & 0xFCu) << 8) + (g & 3u);  // row (g>>2) is 256 words = 1024 bytes long
    *p = (uint8_t)(*p + 1);
}

__device__ __forceinline__ void hp_bump4(uint8_t *mine, uint32_t g4)
{
    hp_bump(mine, g4 & 0xFFu);
    hp_bump(mine, (g4 >> 8) & 0xFFu);
    hp_bump(mine, (g4 >> 16) & 0xFFu);
    hp_bump(mine, g4 >> 24);
}

template <bool STORE>
__global__ void __launch_bounds__(HP_THREADS, 3) gray_hist_private_kernel(const uint4 *__restrict__ src,
                                                                          uint4 *__restrict__ dst, size_t ngroups,
                                                                          size_t npix, uint32_t per_thread,
                                                                          unsigned long long *d_hist)
{
    PDL_PROLOGUE();
    extern __shared__ __align__(16) uint32_t hp_smem[];
    uint32_t *counters = hp_smem, *total = hp_smem + 64 * HP_THREADS;
    const uint32_t tid = threadIdx.x;
    uint8_t *mine = reinterpret_cast<uint8_t *>(counters + tid);
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < 64 * HP_THREADS / 4; i += HP_THREADS) reinterpret_cast<uint4 *>(counters)[i] = zero4;
    total[tid] = 0;
    __syncthreads();

    const size_t chunk = (size_t)per_thread * HP_THREADS;  // groups per CTA between two folds
    for (size_t base = (size_t)blockIdx.x * chunk; base < ngroups; base += (size_t)gridDim.x * chunk) {
        size_t g = base + tid;
        const size_t end = (base + chunk < ngroups) ? base + chunk : ngroups;
        uint4 a, b, c;
        if (g < end) {
            a = __ldg(src + 3 * g);
            b = __ldg(src + 3 * g + 1);
            c = __ldg(src + 3 * g + 2);
        }
        while (g < end) {
            const size_t gn = g + HP_THREADS;
            uint4 na, nb, nc;
            if (gn < end) {  // next group's loads fly while this one is counted
                na = __ldg(src + 3 * gn);
                nb = __ldg(src + 3 * gn + 1);
                nc = __ldg(src + 3 * gn + 2);
            }
            const uint4 o = gray16(a, b, c);
            if (STORE) dst[g] = o;
            hp_bump4(mine, o.x);
            hp_bump4(mine, o.y);
            hp_bump4(mine, o.z);
            hp_bump4(mine, o.w);
            a = na;
            b = nb;
            c = nc;
            g = gn;
        }
        __syncthreads();
        {  // fold: thread `tid` sums bin `tid` over all 256 columns, rows skewed across banks
            const uint32_t row = tid >> 2, sel = 1u << (8u * (tid & 3u));
            const uint4 *rowp = reinterpret_cast<const uint4 *>(counters + row * HP_THREADS);
            uint32_t acc = 0;
#pragma unroll 8
            for (uint32_t k = 0; k < 64; k++) {
                const uint4 v = rowp[(k + row) & 63u];
                acc = __dp4a(v.x, sel, acc);
                acc = __dp4a(v.y, sel, acc);
                acc = __dp4a(v.z, sel, acc);
                acc = __dp4a(v.w, sel, acc);
            }
            total[tid] += acc;
        }
        __syncthreads();
        for (uint32_t i = tid; i < 64 * HP_THREADS / 4; i += HP_THREADS) reinterpret_cast<uint4 *>(counters)[i] = zero4;
        __syncthreads();
    }
    // the last (npix % 16) pixels, scalar, straight into the per-CTA totals
    if (blockIdx.x == 0 && tid < (npix - ngroups * 16)) {
        const size_t i = ngroups * 16 + tid;
        const uint8_t *s8 = reinterpret_cast<const uint8_t *>(src) + 3 * i;
        const uint32_t g = div3((uint32_t)s8[0] + s8[1] + s8[2]);
        if (STORE) reinterpret_cast<uint8_t *>(dst)[i] = (uint8_t)g;
        atomicAdd(&total[g], 1u);
    }
    __syncthreads();
    if (total[tid]) atomicAdd(&d_hist[tid], (unsigned long long)total[tid]);
}

// ---- histogram with one shared-memory column per LANE: bins[256][32] u32.  Lane l of every warp
// only ever touches bank l, so a warp's 32 updates never conflict (a constant image costs the same
// as noise); warps share columns, hence RED.shared adds.  One fold + one global atomic pass per CTA.
constexpr size_t HL_SMEM = 256 * 32 * sizeof(uint32_t);

__device__ __forceinline__ uint32_t hl_gray4(uint32_t *col, uint32_t a, uint32_t b, uint32_t c)
{  // grey of 4 pixels, each counted in this lane's column before the bytes are packed
    uint32_t q[4];
    gray4_split(a, b, c, q);
#pragma unroll
    for (int i = 0; i < 4; i++) atomicAdd(col + q[i] * 32u, 1u);
    return pack4(q);
}

constexpr int HL_THREADS = 1024;
template <bool STORE>
__global__ void __launch_bounds__(HL_THREADS, 1) gray_hist_lanes_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                              size_t ngroups, size_t npix,
                                                              unsigned long long *d_hist)
{
    pdl_trigger();
    extern __shared__ __align__(16) uint32_t hl_bins[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < 256 * 32 / 4; i += HL_THREADS) reinterpret_cast<uint4 *>(hl_bins)[i] = zero4;
    __syncthreads();
    pdl_wait();  // everything above is index arithmetic; global memory is touched only below
    uint32_t *col = hl_bins + lane;
    // few, fat CTAs (one per SM): the final 256 global atomics per CTA hit only 16 cache lines, and
    // every CTA adds to all of them, so the number of CTAs is what that last pass costs
    const size_t stride = (size_t)gridDim.x * HL_THREADS;
    size_t g = (size_t)blockIdx.x * HL_THREADS + tid;
    for (; g + stride < ngroups; g += 2 * stride) {  // two groups (96 B) in flight per thread
        const uint4 *p = src + 3 * g, *p2 = src + 3 * (g + stride);
        const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        const uint4 a2 = __ldg(p2), b2 = __ldg(p2 + 1), c2 = __ldg(p2 + 2);
        uint4 o, o2;
        o.x = hl_gray4(col, a.x, a.y, a.z);
        o.y = hl_gray4(col, a.w, b.x, b.y);
        o.z = hl_gray4(col, b.z, b.w, c.x);
        o.w = hl_gray4(col, c.y, c.z, c.w);
        if (STORE) dst[g] = o;
        o2.x = hl_gray4(col, a2.x, a2.y, a2.z);
        o2.y = hl_gray4(col, a2.w, b2.x, b2.y);
        o2.z = hl_gray4(col, b2.z, b2.w, c2.x);
        o2.w = hl_gray4(col, c2.y, c2.z, c2.w);
        if (STORE) dst[g + stride] = o2;
    }
    if (g < ngroups) {
        const uint4 *p = src + 3 * g;
        const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        uint4 o;
        o.x = hl_gray4(col, a.x, a.y, a.z);
        o.y = hl_gray4(col, a.w, b.x, b.y);
        o.z = hl_gray4(col, b.z, b.w, c.x);
        o.w = hl_gray4(col, c.y, c.z, c.w);
        if (STORE) dst[g] = o;
    }
    if (blockIdx.x == 0 && tid < (npix - ngroups * 16)) {
        const size_t i = ngroups * 16 + tid;
        const uint8_t *s8 = reinterpret_cast<const uint8_t *>(src) + 3 * i;
        const uint32_t g = div3((uint32_t)s8[0] + s8[1] + s8[2]);
        if (STORE) reinterpret_cast<uint8_t *>(dst)[i] = (uint8_t)g;
        atomicAdd(col + (g << 5), 1u);
    }
    __syncthreads();
    if (tid < 256) {  // fold bin `tid` over its 32 columns, starting at a different bank per lane
        const uint32_t *row = hl_bins + tid * 32u;
        uint32_t t = 0;
#pragma unroll
        for (uint32_t k = 0; k < 32; k++) t += row[(k + lane) & 31u];
        if (t) atomicAdd(&d_hist[tid], (unsigned long long)t);
    }
}

// variant: one 16-pixel group per thread, one CTA per 256 groups (no grid-stride loop)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) gray_flat_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                          size_t ngroups, size_t npix)
{
    pdl_trigger();
    const size_t g = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    pdl_wait();  // everything above is index arithmetic; global memory is touched only below
    if (g < ngroups) {
        const uint4 *p = src + 3 * g;
        const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        dst[g] = gray16(a, b, c);
    }
    if (blockIdx.x == 0 && threadIdx.x < (npix - ngroups * 16)) {
        const size_t i = ngroups * 16 + threadIdx.x;
        const uint8_t *s8 = reinterpret_cast<const uint8_t *>(src) + 3 * i;
        reinterpret_cast<uint8_t *>(dst)[i] = (uint8_t)div3((uint32_t)s8[0] + s8[1] + s8[2]);
    }
}

// ---- TMA bulk-copy pipeline (cp.async.bulk + mbarrier): one elected thread streams 12 KB tiles
// of the raster into a ring of shared-memory stages; the CTA reads each tile conflict-free
// (48 B per thread), computes, and stores 16 B per thread straight to global memory.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

constexpr int GT_STAGES = 4;
constexpr int GT_TILE_GROUPS = 256;                 // one 16-pixel group per thread
constexpr int GT_TILE_BYTES = GT_TILE_GROUPS * 48;  // 12288
constexpr size_t GT_SMEM = (size_t)GT_STAGES * GT_TILE_BYTES + 128;

__global__ void __launch_bounds__(256) gray_tma_kernel(const uint8_t *__restrict__ src, uint4 *__restrict__ dst,
                                                       size_t ngroups, size_t npix)
{
    PDL_PROLOGUE();
    extern __shared__ __align__(128) uint8_t gt_smem[];
    __shared__ __align__(8) uint64_t full[GT_STAGES];
    const uint32_t tid = threadIdx.x;
    const size_t ntiles = (ngroups + GT_TILE_GROUPS - 1) / GT_TILE_GROUPS;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t tile, int s) {
        const size_t g0 = tile * GT_TILE_GROUPS;
        const uint32_t groups = (uint32_t)((ngroups - g0 < GT_TILE_GROUPS) ? ngroups - g0 : GT_TILE_GROUPS);
        mbar_expect_tx(&full[s], groups * 48u);
        bulk_g2s(gt_smem + (size_t)s * GT_TILE_BYTES, src + g0 * 48, groups * 48u, &full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < GT_STAGES; s++) {
            const size_t t = (size_t)blockIdx.x + (size_t)s * gridDim.x;
            if (t < ntiles) issue(t, s);
        }
    uint32_t it = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int s = it % GT_STAGES;
        mbar_wait(&full[s], (it / GT_STAGES) & 1u);
        const uint4 *p = reinterpret_cast<const uint4 *>(gt_smem + (size_t)s * GT_TILE_BYTES) + 3 * tid;
        const uint4 a = p[0], b = p[1], c = p[2];
        __syncthreads();  // every thread has read stage s: it may be refilled
        if (tid == 0) {
            const size_t nt = tile + (size_t)GT_STAGES * gridDim.x;
            if (nt < ntiles) issue(nt, s);
        }
        const size_t g = tile * GT_TILE_GROUPS + tid;
        if (g < ngroups) dst[g] = gray16(a, b, c);
    }
    if (blockIdx.x == 0 && tid < (npix - ngroups * 16)) {
        const size_t i = ngroups * 16 + tid;
        const uint8_t *s8 = src + 3 * i;
        reinterpret_cast<uint8_t *>(dst)[i] = (uint8_t)div3((uint32_t)s8[0] + s8[1] + s8[2]);
    }
}

template <bool HIST, bool STORE>
static cudaError_t gray_dispatch(const uint8_t *src, uint8_t *dst, size_t npix, unsigned long long *d_hist,
                                 cudaStream_t s)
{
    if (npix == 0) return cudaSuccess;
    if (aligned16(src) && (!STORE || aligned16(dst))) {
        size_t ngroups = npix / 16;
        if (HIST && g_variant != 2 && g_variant != 5) {
            // lane-private columns, RED.shared; one 1024-thread CTA per SM
            static bool ok[64] = {};
            allow_smem(gray_hist_lanes_kernel<STORE>, HL_SMEM, ok);
            size_t want = (ngroups + HL_THREADS - 1) / HL_THREADS, wave = (size_t)sm_count();
            unsigned grid = (unsigned)(want < 1 ? 1 : want < wave ? want : wave);
            launch(gray_hist_lanes_kernel<STORE>, dim3(grid), dim3(HL_THREADS), HL_SMEM, s, reinterpret_cast<const uint4 *>(src),
                   reinterpret_cast<uint4 *>(dst), ngroups, npix, d_hist);
            return PPMX_LAUNCHED();
        }
        if (HIST && g_variant == 5) {
            // thread-private byte counters: <= 15 groups per thread between folds, 3 CTAs per SM
            static bool ok[64] = {};
            allow_smem(gray_hist_private_kernel<STORE>, HP_SMEM, ok);
            size_t wave = (size_t)sm_count() * 3 * HP_THREADS;
            size_t per = (ngroups + wave - 1) / wave;
            if (per < 1) per = 1;
            if (per > HP_MAX_GROUPS) per = HP_MAX_GROUPS;
            size_t chunks = (ngroups + per * HP_THREADS - 1) / (per * HP_THREADS);
            if (chunks < 1) chunks = 1;
            unsigned grid = (unsigned)(chunks < (size_t)sm_count() * 3 ? chunks : (size_t)sm_count() * 3);
            launch(gray_hist_private_kernel<STORE>, dim3(grid), dim3(HP_THREADS), HP_SMEM, s,
                   reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), ngroups, npix, (uint32_t)per, d_hist);
            return PPMX_LAUNCHED();
        }
        if (!HIST && g_variant != 1 && g_variant != 4) {  // default: one group per thread, no loop
            if (g_variant == 6 || g_variant == 7) {  // smaller CTAs: shorter tail, more CTA launches
                const unsigned blk = g_variant == 6 ? 128u : 64u;
                unsigned grid = (unsigned)((ngroups + blk - 1) / blk);
                if (blk == 128) launch(gray_flat_kernel<128>, dim3(grid ? grid : 1), dim3(128), 0, s,
                                       reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), ngroups, npix);
                else launch(gray_flat_kernel<64>, dim3(grid ? grid : 1), dim3(64), 0, s,
                            reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), ngroups, npix);
                return PPMX_LAUNCHED();
            }
            unsigned grid = (unsigned)((ngroups + 255) / 256);
            launch(gray_flat_kernel<256>, dim3(grid ? grid : 1), dim3(256), 0, s, reinterpret_cast<const uint4 *>(src),
                   reinterpret_cast<uint4 *>(dst), ngroups, npix);
            return PPMX_LAUNCHED();
        }
        if (!HIST && g_variant == 4) {
            static bool ok[64] = {};
            allow_smem(gray_tma_kernel, GT_SMEM, ok);
            size_t ntiles = (ngroups + GT_TILE_GROUPS - 1) / GT_TILE_GROUPS;
            unsigned grid = (unsigned)(ntiles < (size_t)sm_count() * 4 ? (ntiles ? ntiles : 1) : (size_t)sm_count() * 4);
            launch(gray_tma_kernel, dim3(grid), dim3(256), GT_SMEM, s, src, reinterpret_cast<uint4 *>(dst), ngroups, npix);
            return PPMX_LAUNCHED();
        }
        unsigned grid = wave_grid(ngroups ? ngroups : 1, 256, 8);
        launch(gray_vec_kernel<HIST, STORE>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const uint4 *>(src),
                                                          reinterpret_cast<uint4 *>(dst), ngroups, npix, d_hist);
    } else {
        launch(gray_scalar_kernel<HIST, STORE>, dim3(wave_grid(npix, 256, 8)), dim3(256), 0, s, src, dst, npix, d_hist);
    }
    return PPMX_LAUNCHED();
}

cudaError_t gray(const uint8_t *src, uint8_t *dst, size_t npix, unsigned long long *d_hist, cudaStream_t s)
{
    return d_hist ? gray_dispatch<true, true>(src, dst, npix, d_hist, s)
                  : gray_dispatch<false, true>(src, dst, npix, nullptr, s);
}

cudaError_t hist_gray(const uint8_t *src, size_t npix, unsigned long long *d_hist, cudaStream_t s)
{
    return gray_dispatch<true, false>(src, nullptr, npix, d_hist, s);
}

// ------------------------------------------------------------------------------------------
// mono  (ref:964-969), alone (R8 of 0/1) and fused with the P4 packer (ref:268-284)
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) mono_plane_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                         uint32_t w, uint32_t h, uint32_t y0)
{
    PDL_PROLOGUE();
    const size_t n = (size_t)w * h, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (size_t)y * w);
        uint32_t g = div3((uint32_t)src[3 * i] + src[3 * i + 1] + src[3 * i + 2]);
        dst[i] = (g < c_bayer[(x & 3u) * 4 + ((y + y0) & 3u)]) ? 1 : 0;
    }
}

// w % 16 == 0 and 16-byte aligned rasters: one thread = 16 pixels of one row (48 B in) = two
// output bytes; one CTA per 256 such groups, no loop (same access pattern as gray)
// grey < thr  <=>  (r+g+b)/3 < thr  <=>  r+g+b < 3*thr (integers), so neither the division nor the
// grey byte is needed: each dp4a starts from -3*thr and the pixel's bit is the SIGN of the sum, which
// one funnel shift appends to the output (pixel 0 ends up in the most significant bit, ref:273).
__device__ __forceinline__ uint32_t mono_bits4(uint32_t bits, uint32_t a, uint32_t b, uint32_t c, const int (&t)[4])
{
    const int s0 = (int)__dp4a(a, 0x00010101u, (uint32_t)t[0]);
    const int s1 = (int)__dp4a(a, 0x01000000u, __dp4a(b, 0x00000101u, (uint32_t)t[1]));
    const int s2 = (int)__dp4a(b, 0x01010000u, __dp4a(c, 0x00000001u, (uint32_t)t[2]));
    const int s3 = (int)__dp4a(c, 0x01010100u, (uint32_t)t[3]);
    bits = __funnelshift_l((uint32_t)s0, bits, 1);
    bits = __funnelshift_l((uint32_t)s1, bits, 1);
    bits = __funnelshift_l((uint32_t)s2, bits, 1);
    bits = __funnelshift_l((uint32_t)s3, bits, 1);
    return bits;
}

__global__ void __launch_bounds__(256) mono_bits_vec_kernel(const uint4 *__restrict__ src, uint16_t *__restrict__ dst,
                                                            uint32_t groups_per_row, size_t ngroups, uint32_t y0)
{
    pdl_trigger();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= ngroups) return;
    const uint32_t y = (ngroups <= 0xFFFFFFFFull ? (uint32_t)i / groups_per_row : (uint32_t)(i / groups_per_row)) + y0;
    const uint32_t yy = y & 3u;
    const int t[4] = {-3 * (int)c_bayer[yy], -3 * (int)c_bayer[4 + yy], -3 * (int)c_bayer[8 + yy],
                      -3 * (int)c_bayer[12 + yy]};  // x % 4 = 0..3 on this row (a group starts at x % 16 == 0)
    pdl_wait();  // everything above is index arithmetic; global memory is touched only below
    const uint4 *p = src + 3 * i;
    const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    uint32_t bits = 0;
    bits = mono_bits4(bits, a.x, a.y, a.z, t);
    bits = mono_bits4(bits, a.w, b.x, b.y, t);
    bits = mono_bits4(bits, b.z, b.w, c.x, t);
    bits = mono_bits4(bits, c.y, c.z, c.w, t);
    // pixels 0-7 sit in bits 15..8: they are the FIRST byte in memory
    dst[i] = (uint16_t)__byte_perm(bits, 0, 0x4401);
}

// any width / alignment: one thread = one output byte (up to 8 pixels of one row)
__global__ void __launch_bounds__(256) mono_bits_generic_kernel(const uint8_t *__restrict__ src,
                                                                uint8_t *__restrict__ dst, uint32_t w, uint32_t h,
                                                                uint32_t row_bytes, uint32_t y0)
{
    PDL_PROLOGUE();
    const size_t n = (size_t)row_bytes * h, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t y = (uint32_t)(i / row_bytes), bx = (uint32_t)(i - (size_t)y * row_bytes);
        uint32_t x0 = bx * 8, cnt = min(8u, w - x0), yy = (y + y0) & 3u;
        const uint8_t *p = src + ((size_t)y * w + x0) * 3;
        uint32_t out = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            uint32_t g = div3((uint32_t)p[3 * j] + p[3 * j + 1] + p[3 * j + 2]);
            if (g < c_bayer[((x0 + j) & 3u) * 4 + yy]) out |= 0x80u >> j;
        }
        dst[i] = (uint8_t)out;
    }
}

cudaError_t mono_plane(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t y0, cudaStream_t s)
{
    size_t n = (size_t)w * h;
    if (!n) return cudaSuccess;
    launch(mono_plane_kernel, dim3(wave_grid(n, 256, 8)), dim3(256), 0, s, src, dst, w, h, y0);
    return PPMX_LAUNCHED();
}

cudaError_t mono_bits(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t y0, cudaStream_t s)
{
    if (!w || !h) return cudaSuccess;
    if ((w % 16u) == 0 && aligned16(src) && aligned4(dst)) {
        size_t ngroups = (size_t)(w / 16u) * h;
        launch(mono_bits_vec_kernel, dim3((unsigned)((ngroups + 255) / 256)), dim3(256), 0, s,
               reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint16_t *>(dst), w / 16u, ngroups, y0);
    } else {
        uint32_t rb = (w + 7u) / 8u;
        launch(mono_bits_generic_kernel, dim3(wave_grid((size_t)rb * h, 256, 8)), dim3(256), 0, s, src, dst, w, h, rb, y0);
    }
    return PPMX_LAUNCHED();
}

// the P4 writer alone (ref:268-284) on arbitrary .r bytes: byte |= (r << (7 - x%8)) & 0xff
__global__ void __launch_bounds__(256) pack_pbm_kernel(const uint8_t *__restrict__ src, int bpp,
                                                       uint8_t *__restrict__ dst, uint32_t w, uint32_t h,
                                                       uint32_t row_bytes)
{
    PDL_PROLOGUE();
    const size_t n = (size_t)row_bytes * h, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t y = (uint32_t)(i / row_bytes), bx = (uint32_t)(i - (size_t)y * row_bytes);
        uint32_t x0 = bx * 8, cnt = min(8u, w - x0);
        const uint8_t *p = src + ((size_t)y * w + x0) * bpp;
        uint32_t out = 0;
        for (uint32_t j = 0; j < cnt; j++) out |= ((uint32_t)p[(size_t)j * bpp] << (7 - j));
        dst[i] = (uint8_t)(out & 0xFFu);
    }
}

cudaError_t pack_pbm(const uint8_t *src, int src_bpp, uint8_t *dst, uint32_t w, uint32_t h, cudaStream_t s)
{
    if (!w || !h) return cudaSuccess;
    uint32_t rb = (w + 7u) / 8u;
    launch(pack_pbm_kernel, dim3(wave_grid((size_t)rb * h, 256, 8)), dim3(256), 0, s, src, src_bpp, dst, w, h, rb);
    return PPMX_LAUNCHED();
}

// the PGM writer's gather (ref:263-267): .r of every pixel
__global__ void __launch_bounds__(256) extract_r_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                        size_t npix)
{
    PDL_PROLOGUE();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) dst[i] = src[3 * i];
}

cudaError_t extract_r(const uint8_t *src, uint8_t *dst, size_t npix, cudaStream_t s)
{
    if (!npix) return cudaSuccess;
    launch(extract_r_kernel, dim3(wave_grid(npix, 256, 8)), dim3(256), 0, s, src, dst, npix);
    return PPMX_LAUNCHED();
}

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

cudaError_t flip(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int bpp, int vertical, cudaStream_t s)
{
    if (!w || !h) return cudaSuccess;
    size_t row = (size_t)w * bpp;
    if (vertical) {
        if (row % 16 == 0 && aligned16(src) && aligned16(dst)) {
            size_t n = row / 16 * h;
            launch(flipv_kernel<uint4>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s,
                   reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), (uint32_t)(row / 16), h, n);
        } else if (row % 4 == 0 && aligned4(src) && aligned4(dst)) {
            size_t n = row / 4 * h;
            launch(flipv_kernel<uint32_t>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s,
                   reinterpret_cast<const uint32_t *>(src), reinterpret_cast<uint32_t *>(dst), (uint32_t)(row / 4), h, n);
        } else {
            launch(flipv_kernel<uint8_t>, dim3((unsigned)((row * h + 255) / 256)), dim3(256), 0, s, src, dst, (uint32_t)row, h,
                   row * h);
        }
    } else {
        if (bpp == 3 && (w % 16u) == 0 && aligned16(src) && aligned16(dst)) {
            size_t n = (size_t)(w / 16u) * h;
            launch(fliph_rgb16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s,
                   reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), w / 16u, n);
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
constexpr int RT_PITCH = RT * 3 + 4;  // bytes; +4 keeps column reads off a single bank

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

// Fast path (w % 16 == 0, h % 16 == 0, 16-byte aligned rasters): 64 x 64 pixel tiles.
//   phase 1: a thread loads 16 pixels of one source row (3 x 16 B), widens them to one word per
//            pixel (r g b x) and stores 4 x 16 B into a swizzled shared tile (no bank conflicts);
//   phase 2: a lane reads one source column over 16 source rows, one word per row (conflict
//            free), repacks the 16 pixels to 48 B and writes them as three 16-byte stores into
//            the destination row that column became; 4 lanes complete 192 contiguous bytes.
constexpr int XT = 64;

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
        } else {
            launch(reverse_pixels_kernel, dim3(wave_grid(npix * 3, 256, 8)), dim3(256), 0, s, src, dst, npix);
        }
        return PPMX_LAUNCHED();
    }
    if (angle != 90 && angle != 270) return cudaErrorInvalidValue;
    if ((w % 16u) == 0 && (h % 16u) == 0 && aligned16(src) && aligned16(dst) && g_variant != 1) {
        // (numbering the CTAs down bands of 2..16 tile rows, for DRAM page locality on the write side,
        // measured 1-5 % SLOWER than this plain 2-D grid: the index arithmetic costs more than it gains)
        // two tiles per CTA at <= 40 registers (6 CTAs per SM) measured best: 0.725 / 0.828 of the HBM
        // roofline at 4096^2 / 16384^2; one tile per CTA (variant 6): 0.723 / 0.753
        const int nt = (g_variant == 6) ? 1 : 2;
        dim3 g64((w + XT * nt - 1) / (XT * nt), (h + XT - 1) / XT);
        if (g64.y > 65535u) return cudaErrorInvalidValue;
        if (g_variant == 6) {
            if (angle == 90) launch(rotate_transpose64_kernel<true, 1, 1, 0>, dim3(g64), dim3(256), 0, s, src, dst, w, h);
            else launch(rotate_transpose64_kernel<false, 1, 1, 0>, dim3(g64), dim3(256), 0, s, src, dst, w, h);
        } else if (g_variant == 7 || g_variant == 8) {
            const unsigned band = g_variant == 7 ? 8u : 4u;
            dim3 gb(g64.x * band, (g64.y + band - 1) / band);
            if (g_variant == 7) {
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
    dim3 grid((w + RT - 1) / RT, (h + RT - 1) / RT);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
    if (angle == 90) launch(rotate_transpose_kernel<true>, dim3(grid), dim3(256), 0, s, src, dst, w, h);
    else if (angle == 270) launch(rotate_transpose_kernel<false>, dim3(grid), dim3(256), 0, s, src, dst, w, h);
    else return cudaErrorInvalidValue;
    return PPMX_LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// FP64 helpers: every operation is a separately rounded IEEE multiply or add, in the
// reference's order; nvcc may not contract them (intrinsics) and the file is built -fmad=false.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// exact u8 -> double without the slow I2F.F64 path: 2^52 + v has v in its low mantissa bits
__device__ __forceinline__ double u8_to_double(uint32_t v)
{
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

// four u8 -> double conversions from one word.
//   CONV 0: I2F.F64.U8 with a byte selector (cvt.rn.f64.u8 of the shifted word folds into it);
//           XU pipe, measured 15.2 conversions/clk/SM (tools/dp_peak).
//   CONV 1: the exact 2^52 trick: PRMT + one DADD on the FP64 pipe (64 inst/clk/SM).
//   CONV 2: bytes 0 and 2 on the XU pipe, bytes 1 and 3 on the FP64 pipe, so neither pipe alone
//           limits the K-tap loops (XU 16/clk vs FP64 64/clk at 2-3 DP instructions per tap byte).
__device__ __forceinline__ double cvt_byte_xu(uint32_t shifted)
{
    double d;
    asm("cvt.rn.f64.u8 %0, %1;" : "=d"(d) : "r"(shifted));
    return d;
}
__device__ __forceinline__ double cvt_byte_dp(uint32_t w, int i)
{
    return __hiloint2double(0x43300000, (int)__byte_perm(w, 0, 0x4440 | i)) - 4503599627370496.0;
}
template <int CONV>
__device__ __forceinline__ void word_to_double4(uint32_t w, double (&d)[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool xu = (CONV == 0) || (CONV == 2 && (i & 1) == 0);
        d[i] = xu ? cvt_byte_xu(w >> (8 * i)) : cvt_byte_dp(w, i);
    }
}

// Keys cubic convolution kernel, a = -0.5 (ref:477-489), same association as the source
__device__ __forceinline__ double cubic(double x)
{
    double a1 = fabs(x), a2 = dmul(a1, a1), a3 = dmul(a2, a1), r = 0.0;
    if (a1 <= 1.0) r = dadd(dsub(dmul(1.5, a3), dmul(2.5, a2)), 1.0);
    if (1.0 < a1 && a1 <= 2.0) {
        double t = dadd(dmul(-0.5, a3), dmul(2.5, a2));
        t = dsub(t, dmul(4.0, a1));
        t = dadd(t, 2.0);
        r = dadd(r, t);
    }
    return r;
}

__device__ __forceinline__ double round_half_up(double v) { return floor(dadd(v, 0.5)); }  // ref:27

// ------------------------------------------------------------------------------------------
// rotate, arbitrary angle  (ref:726-786): inverse map + 4x4 bicubic, nearest on a 2-pixel ring
// ------------------------------------------------------------------------------------------

// WORDS: w % 4 == 0 and an aligned raster -- the 12 bytes of a tap row come in as aligned words.
template <bool WORDS, int CONV>
__global__ void __launch_bounds__(256) rotate_bicubic_kernel(const uint8_t *__restrict__ src,
                                                             uint8_t *__restrict__ dst, uint32_t w, uint32_t h,
                                                             uint32_t nw, uint32_t nh, double cs, double sn,
                                                             int xc, int yc, int xo, int yo)
{
    PDL_PROLOGUE();
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nw || y >= nh) return;
    uint8_t *out = dst + ((size_t)y * nw + x) * 3;

    const int x0 = ((int)x - xo) - xc, y0 = ((int)y - yo) - yc;                              // ref:731-735
    const double nX = dadd(dadd(dmul(cs, (double)x0), dmul(sn, (double)y0)), (double)xc);     // ref:741
    const double nY = dadd(dadd(-dmul(sn, (double)x0), dmul(cs, (double)y0)), (double)yc);    // ref:742
    const double rx = round_half_up(nX), ry = round_half_up(nY);

    uint32_t r = 0, g = 0, b = 0;  // uncovered output stays 0 (ref:727)
    if (rx < (double)w && ry < (double)h && ry >= 0.0 && rx >= 0.0) {                         // ref:744
        if (rx > 1.0 && ry > 1.0 && rx < (double)(uint32_t)(w - 2u) && ry < (double)(uint32_t)(h - 2u)) {  // ref:752
            const double fx = floor(nX), fy = floor(nY);
            double wx[4], wy[4];
            int u0 = (int)(fx - 1.0), v0 = (int)(fy - 1.0);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int u = (int)dadd(dsub(fx, 1.0), (double)i);  // ref:761
                int v = (int)dadd(dsub(fy, 1.0), (double)i);  // ref:758
                wx[i] = cubic(dsub(nX, (double)u));
                wy[i] = cubic(dsub(nY, (double)v));
            }
            double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double p0 = 0.0, p1 = 0.0, p2 = 0.0;
                if (WORDS) {
                    const uint32_t b0 = 3u * (uint32_t)u0, sh = (b0 & 3u) * 8u;
                    const uint32_t *rw = reinterpret_cast<const uint32_t *>(src + (size_t)(v0 + j) * w * 3) + (b0 >> 2);
                    const uint32_t q0w = __ldg(rw), q1w = __ldg(rw + 1), q2w = __ldg(rw + 2);
                    const uint32_t q3w = (b0 & 3u) ? __ldg(rw + 3) : 0u;  // only needed when the run is unaligned
                    double d[3][4];
                    word_to_double4<CONV>(__funnelshift_r(q0w, q1w, sh), d[0]);
                    word_to_double4<CONV>(__funnelshift_r(q1w, q2w, sh), d[1]);
                    word_to_double4<CONV>(__funnelshift_r(q2w, q3w, sh), d[2]);
#pragma unroll
                    for (int i = 0; i < 4; i++) {  // ref:762-764
                        p0 = dadd(p0, dmul(d[(3 * i) >> 2][(3 * i) & 3], wx[i]));
                        p1 = dadd(p1, dmul(d[(3 * i + 1) >> 2][(3 * i + 1) & 3], wx[i]));
                        p2 = dadd(p2, dmul(d[(3 * i + 2) >> 2][(3 * i + 2) & 3], wx[i]));
                    }
                } else {
                    const uint8_t *row = src + ((size_t)(v0 + j) * w + u0) * 3;
#pragma unroll
                    for (int i = 0; i < 4; i++) {  // ref:762-764
                        p0 = dadd(p0, dmul(u8_to_double(row[3 * i]), wx[i]));
                        p1 = dadd(p1, dmul(u8_to_double(row[3 * i + 1]), wx[i]));
                        p2 = dadd(p2, dmul(u8_to_double(row[3 * i + 2]), wx[i]));
                    }
                }
                q0 = dadd(q0, dmul(p0, wy[j]));  // ref:766-768
                q1 = dadd(q1, dmul(p1, wy[j]));
                q2 = dadd(q2, dmul(p2, wy[j]));
            }
            if (q0 < 0.0) q0 = 0.0;  // ref:771-777
            if (q1 < 0.0) q1 = 0.0;
            if (q2 < 0.0) q2 = 0.0;
            if (q0 >= 256.0) q0 = 255.0;
            if (q1 >= 256.0) q1 = 255.0;
            if (q2 >= 256.0) q2 = 255.0;
            // truncation (ref:779-781) of a value in [0, 256): floor, as the low word of q + 1.5*2^52
            r = (uint32_t)__double2loint(__dadd_rd(q0, 6755399441055744.0));
            g = (uint32_t)__double2loint(__dadd_rd(q1, 6755399441055744.0));
            b = (uint32_t)__double2loint(__dadd_rd(q2, 6755399441055744.0));
        } else {  // nearest, ref:783
            const uint8_t *p = src + ((size_t)(int)ry * w + (size_t)(int)rx) * 3;
            r = p[0];
            g = p[1];
            b = p[2];
        }
    }
    out[0] = (uint8_t)r;
    out[1] = (uint8_t)g;
    out[2] = (uint8_t)b;
}

cudaError_t rotate_bicubic(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh,
                           double cos_t, double sin_t, cudaStream_t s)
{
    if (!nw || !nh) return cudaSuccess;
    // centre and offset exactly as ref:694-698 (integer halves)
    int xc = (int)(w / 2u), yc = (int)(h / 2u);
    int xo = (int)(nw / 2u) - (int)(w / 2u), yo = (int)(nh / 2u) - (int)(h / 2u);
    dim3 block(32, 8), grid((nw + 31) / 32, (nh + 7) / 8);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
    if ((w % 4u) == 0 && aligned4(src) && g_variant != 1) {
        if (g_variant == 2)
            launch(rotate_bicubic_kernel<true, 1>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo);
        else if (g_variant == 3)
            launch(rotate_bicubic_kernel<true, 0>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo);
        else
            launch(rotate_bicubic_kernel<true, 2>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo);
    } else {
        launch(rotate_bicubic_kernel<false, 1>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo);
    }
    return PPMX_LAUNCHED();
}

// ---- rows of a raster that may be spread over a band and its neighbours' halos ----------------
__host__ __device__ __forceinline__ int mirror_index(int i, int n)
{
    int m = i % (2 * n);
    if (m < 0) m += 2 * n;
    return m < n ? m : 2 * n - 1 - m;
}

// resolves a row of the WHOLE raster to memory: own band, halo above, halo below (maybe peer HBM)
struct RowSource {
    const uint8_t *own, *top, *bottom;
    int y0, h, halo, full_h;
    __device__ __forceinline__ const uint8_t *row(int gy, size_t pitch) const
    {
        gy = mirror_index(gy, full_h);
        if (gy >= y0 && gy < y0 + h) return own + (size_t)(gy - y0) * pitch;
        if (gy < y0) return top + (size_t)(gy - (y0 - halo)) * pitch;
        return bottom + (size_t)(gy - (y0 + h)) * pitch;
    }
    // the same without the mirror step, for row numbers that are already inside the raster
    __device__ __forceinline__ const uint8_t *row_plain(int gy, size_t pitch) const
    {
        if (gy >= y0 && gy < y0 + h) return own + (size_t)(gy - y0) * pitch;
        if (gy < y0) return top + (size_t)(gy - (y0 - halo)) * pitch;
        return bottom + (size_t)(gy - (y0 + h)) * pitch;
    }
};

static RowSource make_row_source(const uint8_t *src, uint32_t h, const Band &band)
{
    RowSource rs;
    rs.own = src;
    rs.top = band.top;
    rs.bottom = band.bottom;
    rs.y0 = band.full_h ? (int)band.y0 : 0;
    rs.h = (int)h;
    rs.halo = (int)band.halo;
    rs.full_h = band.full_h ? (int)band.full_h : (int)h;
    return rs;
}



// ------------------------------------------------------------------------------------------
// imresize  (ref:820-838 height pass, ref:846-868 width pass): K-tap gather, FP64 accumulate
// in tap order, floor(s + 0.5), clamp, u8 store.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t quantise(double s)
{
    s = round_half_up(s);                                           // ref:831
    return (s < 0.0) ? 0u : (s >= 256.0) ? 255u : (uint32_t)__double2int_rz(s);  // ref:835
}

// The same result with one DP add instead of FRND + compares + F2I: for |v| < 2^31,
// v + 1.5*2^52 rounded toward -inf is floor(v) + 1.5*2^52 exactly and its low word is floor(v) as
// an int32; "< 0 -> 0" and ">= 256 -> 255" on floor(v) (ref:835) become an integer clamp.
__device__ __forceinline__ uint32_t quantise_fast(double s)
{
    const double v = dadd(s, 0.5);                                                // ref:27, 831
    const int n = __double2loint(__dadd_rd(v, 6755399441055744.0));
    return (uint32_t)min(max(n, 0), 255);
}

// height pass, fast path (row pitch % 16 == 0, aligned, K <= 64): one thread = 16 bytes of an
// output row.  The row's K weights and source-row numbers are staged once per CTA in shared memory
// so the K source loads of a thread are independent of each other and fly four at a time.
constexpr int ROWS16_MAXK = 64;
template <int CONV>
__global__ void __launch_bounds__(256) imresize_rows16_kernel(const RowSource src, uint8_t *__restrict__ dst,
                                                              uint32_t row_vecs, int taps,
                                                              const double *__restrict__ wts, const int *__restrict__ idx)
{
    PDL_PROLOGUE();
    __shared__ double s_w[ROWS16_MAXK];
    __shared__ int s_i[ROWS16_MAXK];
    const int y = blockIdx.y;
    if ((int)threadIdx.x < taps) {
        s_w[threadIdx.x] = __ldg(wts + (size_t)y * taps + threadIdx.x);
        s_i[threadIdx.x] = __ldg(idx + (size_t)y * taps + threadIdx.x);
    }
    __syncthreads();
    const uint32_t xv = blockIdx.x * 256 + threadIdx.x;
    if (xv >= row_vecs) return;
    const size_t row_bytes = (size_t)row_vecs * 16;

    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0.0;
    for (int z0 = 0; z0 < taps; z0 += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (z0 + u < taps) v[u] = __ldg(reinterpret_cast<const uint4 *>(src.row_plain(s_i[z0 + u], row_bytes)) + xv);
#pragma unroll
        for (int u = 0; u < 4; u++) {  // tap order is the reference's summation order (ref:826-830)
            if (z0 + u >= taps) break;
            const double wz = s_w[z0 + u];
            const uint32_t wd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double d[4];
                word_to_double4<CONV>(wd[q], d);
#pragma unroll
                for (int b = 0; b < 4; b++) acc[4 * q + b] = dadd(acc[4 * q + b], dmul(d[b], wz));
            }
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
        o[q] = quantise_fast(acc[4 * q]) | (quantise_fast(acc[4 * q + 1]) << 8) | (quantise_fast(acc[4 * q + 2]) << 16) |
               (quantise_fast(acc[4 * q + 3]) << 24);
    reinterpret_cast<uint4 *>(dst + (size_t)y * row_bytes)[xv] = make_uint4(o[0], o[1], o[2], o[3]);
}

// width pass, fast path (w % 4 == 0, aligned, 4 <= K <= 8): one thread = one output column for a
// run of rows; its K weights and indices live in registers.  When the K taps are consecutive source
// pixels (always, except where the table mirrors at the raster's edge) their 3K bytes are fetched as
// aligned words and funnel-shifted into place; otherwise tap by tap.
template <int K, int CONV>
__global__ void __launch_bounds__(128) imresize_colsK_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                             uint32_t w, uint32_t h, int out_w, int rows_per_cta,
                                                             const double *__restrict__ wts, const int *__restrict__ idx)
{
    PDL_PROLOGUE();
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= out_w) return;
    double wk[K];
    int ik[K];
#pragma unroll
    for (int z = 0; z < K; z++) {
        wk[z] = __ldg(wts + (size_t)x * K + z);
        ik[z] = __ldg(idx + (size_t)x * K + z);
    }
    bool consecutive = true;
#pragma unroll
    for (int z = 1; z < K; z++) consecutive = consecutive && (ik[z] == ik[0] + z);
    constexpr int NS = (3 * K + 3) / 4;  // words of the aligned 3K-byte stream
    const uint32_t b0 = 3u * (uint32_t)ik[0], w0 = b0 >> 2, sh = (b0 & 3u) * 8u;
    const uint32_t y0 = blockIdx.y * (uint32_t)rows_per_cta, y1 = min(h, y0 + (uint32_t)rows_per_cta);
    const size_t in_pitch = (size_t)w * 3, out_pitch = (size_t)out_w * 3;
    if (consecutive) {
        // the words of row y+1 are requested before row y is evaluated
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(src + (size_t)y0 * in_pitch) + w0;
        const size_t pitch_words = in_pitch / 4;
        uint32_t q[NS + 1], qn[NS + 1];
#pragma unroll
        for (int j = 0; j <= NS; j++)  // word j is needed iff it starts before the last tap byte
            q[j] = (y0 < y1 && 4u * j < (b0 & 3u) + 3u * K) ? __ldg(rw + j) : 0u;
        for (uint32_t y = y0; y < y1; y++) {
            rw += pitch_words;
#pragma unroll
            for (int j = 0; j <= NS; j++) qn[j] = (y + 1 < y1 && 4u * j < (b0 & 3u) + 3u * K) ? __ldg(rw + j) : 0u;
            double d[NS][4];
#pragma unroll
            for (int j = 0; j < NS; j++) word_to_double4<CONV>(__funnelshift_r(q[j], q[j + 1], sh), d[j]);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int z = 0; z < K; z++) {  // ref:852-858, tap order
                s0 = dadd(s0, dmul(d[(3 * z) >> 2][(3 * z) & 3], wk[z]));
                s1 = dadd(s1, dmul(d[(3 * z + 1) >> 2][(3 * z + 1) & 3], wk[z]));
                s2 = dadd(s2, dmul(d[(3 * z + 2) >> 2][(3 * z + 2) & 3], wk[z]));
            }
            uint8_t *o = dst + (size_t)y * out_pitch + (size_t)x * 3;
            o[0] = (uint8_t)quantise_fast(s0);
            o[1] = (uint8_t)quantise_fast(s1);
            o[2] = (uint8_t)quantise_fast(s2);
#pragma unroll
            for (int j = 0; j <= NS; j++) q[j] = qn[j];
        }
    } else {
        for (uint32_t y = y0; y < y1; y++) {
            const uint8_t *row = src + (size_t)y * in_pitch;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int z = 0; z < K; z++) {
                const uint8_t *p = row + (size_t)ik[z] * 3;
                s0 = dadd(s0, dmul(u8_to_double(p[0]), wk[z]));
                s1 = dadd(s1, dmul(u8_to_double(p[1]), wk[z]));
                s2 = dadd(s2, dmul(u8_to_double(p[2]), wk[z]));
            }
            uint8_t *o = dst + (size_t)y * out_pitch + (size_t)x * 3;
            o[0] = (uint8_t)quantise_fast(s0);
            o[1] = (uint8_t)quantise_fast(s1);
            o[2] = (uint8_t)quantise_fast(s2);
        }
    }
}

template <int K>
static void launch_colsK(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int out_w, const double *wts,
                         const int *idx, cudaStream_t s)
{
    const int rows_per_cta = 16;
    for (uint32_t y0 = 0; y0 < h; y0 += 65535u * rows_per_cta) {
        uint32_t rows = min(65535u * rows_per_cta, h - y0);
        dim3 grid((out_w + 127) / 128, (rows + rows_per_cta - 1) / rows_per_cta);
        if (g_variant == 2)
            launch(imresize_colsK_kernel<K, 1>, grid, dim3(128), 0, s, src + (size_t)y0 * w * 3,
                   dst + (size_t)y0 * out_w * 3, w, rows, out_w, rows_per_cta, wts, idx);
        else if (g_variant == 3)
            launch(imresize_colsK_kernel<K, 0>, grid, dim3(128), 0, s, src + (size_t)y0 * w * 3,
                   dst + (size_t)y0 * out_w * 3, w, rows, out_w, rows_per_cta, wts, idx);
        else
            launch(imresize_colsK_kernel<K, 2>, grid, dim3(128), 0, s, src + (size_t)y0 * w * 3,
                   dst + (size_t)y0 * out_w * 3, w, rows, out_w, rows_per_cta, wts, idx);
    }
}

// height pass: every byte of an output row uses the same K source rows and weights, so the
// raster is treated as rows of 3*w independent bytes; VEC bytes per thread.
template <int VEC>
__global__ void __launch_bounds__(256) imresize_rows_kernel(const RowSource src, uint8_t *__restrict__ dst,
                                                            uint32_t row_bytes, int out_h, int taps,
                                                            const double *__restrict__ wts, const int *__restrict__ idx)
{
    PDL_PROLOGUE();
    const uint32_t xb = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int y = blockIdx.y;
    if (xb >= row_bytes || y >= out_h) return;
    const double *wy = wts + (size_t)y * taps;
    const int *iy = idx + (size_t)y * taps;
    double acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = 0.0;
    for (int z = 0; z < taps; z++) {
        const double wz = __ldg(wy + z);
        const uint8_t *p = src.row_plain(__ldg(iy + z), row_bytes) + xb;
        if (VEC == 4) {
            uint32_t v4 = __ldg(reinterpret_cast<const uint32_t *>(p));
#pragma unroll
            for (int v = 0; v < 4; v++) acc[v] = dadd(acc[v], dmul(u8_to_double((v4 >> (8 * v)) & 0xFFu), wz));
        } else {
            acc[0] = dadd(acc[0], dmul(u8_to_double(p[0]), wz));
        }
    }
    uint8_t *o = dst + (size_t)y * row_bytes + xb;
    if (VEC == 4) {
        uint32_t v4 = quantise(acc[0]) | (quantise(acc[1]) << 8) | (quantise(acc[2]) << 16) | (quantise(acc[3]) << 24);
        *reinterpret_cast<uint32_t *>(o) = v4;
    } else {
        o[0] = (uint8_t)quantise(acc[0]);
    }
}

// width pass: one thread = one output pixel; taps gather 3-byte pixels along the row
__global__ void __launch_bounds__(256) imresize_cols_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                            uint32_t w, uint32_t h, int out_w, int taps,
                                                            const double *__restrict__ wts, const int *__restrict__ idx)
{
    PDL_PROLOGUE();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out_w || y >= h) return;
    const double *wx = wts + (size_t)x * taps;
    const int *ix = idx + (size_t)x * taps;
    const uint8_t *row = src + (size_t)y * w * 3;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int z = 0; z < taps; z++) {
        const double wz = __ldg(wx + z);
        const uint8_t *p = row + (size_t)__ldg(ix + z) * 3;
        s0 = dadd(s0, dmul(u8_to_double(p[0]), wz));
        s1 = dadd(s1, dmul(u8_to_double(p[1]), wz));
        s2 = dadd(s2, dmul(u8_to_double(p[2]), wz));
    }
    uint8_t *o = dst + ((size_t)y * out_w + x) * 3;
    o[0] = (uint8_t)quantise(s0);
    o[1] = (uint8_t)quantise(s1);
    o[2] = (uint8_t)quantise(s2);
}

cudaError_t imresize(const uint8_t *src_ptr, uint8_t *dst, uint32_t w, uint32_t h, int out_size, int dim, int taps,
                     const double *d_weights, const int *d_indices, const Band &band, cudaStream_t s)
{
    if (out_size <= 0 || !w || !h) return cudaSuccess;
    if (dim == 0) {
        // a band computes output rows [out_y0, out_y0 + out_rows) from its own source rows plus halos
        const RowSource src = make_row_source(src_ptr, h, band);
        if (band.full_h) {
            if (band.out_y0 + band.out_rows > (uint32_t)out_size) return cudaErrorInvalidValue;
            d_weights += (size_t)band.out_y0 * taps;
            d_indices += (size_t)band.out_y0 * taps;
            out_size = (int)band.out_rows;
            if (out_size <= 0) return cudaSuccess;
        }
        const bool halo_ok = (!band.top || aligned16(band.top)) && (!band.bottom || aligned16(band.bottom));
        uint32_t row_bytes = w * 3u;
        if (row_bytes % 16 == 0 && aligned16(src_ptr) && aligned16(dst) && halo_ok && taps <= ROWS16_MAXK && g_variant != 1) {
            dim3 grid((row_bytes / 16 + 255) / 256, 1);
            for (int y0 = 0; y0 < out_size; y0 += 65535) {
                int rows = min(65535, out_size - y0);
                grid.y = rows;
                if (g_variant == 2)
                    launch(imresize_rows16_kernel<1>, grid, dim3(256), 0, s, src, dst + (size_t)y0 * row_bytes, row_bytes / 16,
                           taps, d_weights + (size_t)y0 * taps, d_indices + (size_t)y0 * taps);
                else if (g_variant == 3)
                    launch(imresize_rows16_kernel<0>, grid, dim3(256), 0, s, src, dst + (size_t)y0 * row_bytes, row_bytes / 16,
                           taps, d_weights + (size_t)y0 * taps, d_indices + (size_t)y0 * taps);
                else
                    launch(imresize_rows16_kernel<2>, grid, dim3(256), 0, s, src, dst + (size_t)y0 * row_bytes, row_bytes / 16,
                           taps, d_weights + (size_t)y0 * taps, d_indices + (size_t)y0 * taps);
            }
            return cudaGetLastError();
        }
        if (row_bytes % 4 == 0 && aligned4(src_ptr) && aligned4(dst) && halo_ok) {
            dim3 grid((row_bytes / 4 + 255) / 256, 1);
            // rows go on grid.y in slabs of <= 65535
            for (int y0 = 0; y0 < out_size; y0 += 65535) {
                int rows = min(65535, out_size - y0);
                grid.y = rows;
                launch(imresize_rows_kernel<4>, dim3(grid), dim3(256), 0, s, src, dst + (size_t)y0 * row_bytes, row_bytes, rows, taps,
                                                             d_weights + (size_t)y0 * taps, d_indices + (size_t)y0 * taps);
                            }
        } else {
            dim3 grid((row_bytes + 255) / 256, 1);
            for (int y0 = 0; y0 < out_size; y0 += 65535) {
                int rows = min(65535, out_size - y0);
                grid.y = rows;
                launch(imresize_rows_kernel<1>, dim3(grid), dim3(256), 0, s, src, dst + (size_t)y0 * row_bytes, row_bytes, rows, taps,
                                                             d_weights + (size_t)y0 * taps, d_indices + (size_t)y0 * taps);
                            }
        }
        return cudaGetLastError();
    }
    const uint8_t *src = src_ptr;
    if ((w % 4u) == 0 && aligned4(src) && taps >= 4 && taps <= 8 && g_variant != 1) {
        switch (taps) {
        case 4: launch_colsK<4>(src, dst, w, h, out_size, d_weights, d_indices, s); break;
        case 5: launch_colsK<5>(src, dst, w, h, out_size, d_weights, d_indices, s); break;
        case 6: launch_colsK<6>(src, dst, w, h, out_size, d_weights, d_indices, s); break;
        case 7: launch_colsK<7>(src, dst, w, h, out_size, d_weights, d_indices, s); break;
        default: launch_colsK<8>(src, dst, w, h, out_size, d_weights, d_indices, s); break;
        }
        return cudaGetLastError();
    }
    dim3 block(64, 4), grid((out_size + 63) / 64, (h + 3) / 4);
    if (grid.y > 65535u) {
        for (uint32_t y0 = 0; y0 < h; y0 += 65535u * 4u) {
            uint32_t rows = min(65535u * 4u, h - y0);
            dim3 g2(grid.x, (rows + 3) / 4);
            launch(imresize_cols_kernel, dim3(g2), dim3(block), 0, s, src + (size_t)y0 * w * 3, dst + (size_t)y0 * out_size * 3, w, rows,
                                                     out_size, taps, d_weights, d_indices);
                    }
        return cudaGetLastError();
    }
    launch(imresize_cols_kernel, dim3(grid), dim3(block), 0, s, src, dst, w, h, out_size, taps, d_weights, d_indices);
    return PPMX_LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// EXTENSION (no reference counterpart, parity unpinned): k x k integer convolution.
// Border = symmetric mirror (the aux-table idiom of ref:551-555,589); result =
// floor(acc/div + 0.5) + bias computed in integers, clamped like ref:835.  Integer sums are
// exact in any order, so the fast kernel may regroup taps freely and still match the self-oracle.
// ------------------------------------------------------------------------------------------

constexpr int CONV_MAXK = 15;

// exact floor((2*acc + div) / (2*div)) + bias with one multiply-high: the numerator is shifted to
// be non-negative by a multiple K of the divisor d = 2*div, then n/d = (n * M) >> (31 + l) for all
// n < 2^31 with l = ceil(log2 d), M = ceil(2^(31+l) / d)  (Granlund & Montgomery, N = 31).
struct ConvRound {
    uint32_t M, shift;  // shift = l - 1, applied to the high word of n * M
    int32_t add, K, bias;  // n = 2*acc + add, add = div + d*K
    int32_t pow2;          // d is a power of two: a plain shift by l replaces the multiply-high
    __device__ __forceinline__ int32_t quotient(int32_t acc) const  // before the 0..255 clamp
    {
        uint32_t n = (uint32_t)(2 * acc + add);
        uint32_t qn = pow2 ? (n >> (shift + 1)) : (__umulhi(n, M) >> shift);
        return (int32_t)qn - K + bias;
    }
    __device__ __forceinline__ uint32_t operator()(int32_t acc) const { return (uint32_t)min(max(quotient(acc), 0), 255); }
    // four results clamped to 0..255 and packed, result 0 in the low byte: two I2IP instructions
    __device__ __forceinline__ uint32_t pack4(int32_t a0, int32_t a1, int32_t a2, int32_t a3) const
    {
        uint32_t hi, out;
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(quotient(a3)), "r"(quotient(a2)), "r"(0));
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(quotient(a1)), "r"(quotient(a0)), "r"(hi));
        return out;
    }
};

static bool make_conv_round(int64_t sum_abs, int32_t div, int32_t bias, ConvRound *r)
{
    const uint64_t d = 2ull * (uint64_t)div;
    int l = 0;
    while ((1ull << l) < d) l++;
    if (l < 1 || l > 30) return false;
    const uint64_t K = (2ull * 255ull * (uint64_t)sum_abs + d - 1) / d + 1;  // makes every numerator >= 0
    const uint64_t nmax = 2ull * 255ull * (uint64_t)sum_abs + (uint64_t)div + d * K;
    if (nmax >= (1ull << 31) || K >= (1ull << 30)) return false;
    const unsigned __int128 pw = (unsigned __int128)1 << (31 + l);
    r->M = (uint32_t)((pw + d - 1) / d);
    r->shift = (uint32_t)(l - 1);
    r->add = (int32_t)((uint64_t)div + d * K);
    r->K = (int32_t)K;
    r->bias = bias;
    r->pow2 = ((d & (d - 1)) == 0) ? 1 : 0;
    return true;
}

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
constexpr int FC_TW = 128, FC_TH = 32, FC_RV = 4;
constexpr int FC_PITCH = FC_TW + 32;  // one 16-pixel group of halo on each side

// unsigned pixel bytes times signed coefficient bytes (the CUDA intrinsic has no mixed form)
__device__ __forceinline__ int32_t dp4a_u8s8(uint32_t px4, uint32_t coef4, int32_t acc)
{
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px4), "r"(coef4), "r"(acc));
    return d;
}

template <int K>
struct ConvCoefPacked {
    uint32_t cw[K][4][3];
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

template <int K>
__global__ void __launch_bounds__(256) conv_dp4a_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t w,
                                                        const ConvCoefPacked<K> cf, const ConvRound rnd)
{
    PDL_PROLOGUE();
    constexpr int R = K / 2, IN_ROWS = FC_TH + K - 1;
    __shared__ __align__(16) uint8_t plane[3][IN_ROWS][FC_PITCH];
    const int tx0 = blockIdx.x * FC_TW, ty0 = blockIdx.y * FC_TH;
    const size_t pitch = (size_t)w * 3;

    // ---- stage: 16-pixel groups, IN_ROWS x (FC_PITCH / 16) of them ----
    constexpr int GROUPS_X = FC_PITCH / 16;
    for (int i = threadIdx.x; i < IN_ROWS * GROUPS_X; i += 256) {
        const int ty = i / GROUPS_X, gx = i - ty * GROUPS_X;
        const int x0 = tx0 - 16 + 16 * gx;  // first source pixel of the group (may be outside the raster)
        const uint8_t *row = rs.row(rs.y0 + ty0 + ty - R, pitch);
        if (x0 >= 0 && x0 + 16 <= (int)w) {
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
            for (int j = 0; j < 4; j++) acc[a][j] = 0;
#pragma unroll
        for (int iy = 0; iy < FC_RV + K - 1; iy++) {
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(&plane[ch][oy0 + iy][0]) + 3 + lx;
            const uint32_t wsrc[3] = {rw[0], rw[1], rw[2]};  // pixels x-4..x-1, x..x+3, x+4..x+7
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
#pragma unroll
        for (int a = 0; a < FC_RV; a++)
            outw[a][ch] = rnd.pack4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
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

template <int K>
static cudaError_t conv_fast(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef,
                             const ConvRound &rnd, cudaStream_t s)
{
    ConvCoefPacked<K> cf;
    for (int dy = 0; dy < K; dy++)
        for (int j = 0; j < 4; j++)
            for (int wi = 0; wi < 3; wi++) {
                uint32_t word = 0;
                for (int b = 0; b < 4; b++) {
                    int dx = 4 * (wi - 1) + b - j;
                    if (dx >= -(K / 2) && dx <= K / 2)
                        word |= (uint32_t)(uint8_t)(int8_t)coef[dy * K + dx + K / 2] << (8 * b);
                }
                cf.cw[dy][j][wi] = word;
            }
    dim3 grid((w + FC_TW - 1) / FC_TW, (h + FC_TH - 1) / FC_TH);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
    launch(conv_dp4a_kernel<K>, grid, dim3(256), 0, s, rs, dst, w, cf, rnd);
    return PPMX_LAUNCHED();
}

cudaError_t conv(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div,
                 int32_t bias, const Band &band, cudaStream_t s)
{
    if (k < 1 || k > CONV_MAXK || !(k & 1) || div < 1) return cudaErrorInvalidValue;
    if (!w || !h) return cudaSuccess;
    const RowSource rs = make_row_source(src, h, band);

    bool s8 = true;
    int64_t sum_abs = 0;
    for (int i = 0; i < k * k; i++) {
        if (coef[i] < -128 || coef[i] > 127) s8 = false;
        sum_abs += coef[i] < 0 ? -(int64_t)coef[i] : coef[i];
    }
    ConvRound rnd;
    if (s8 && (k == 3 || k == 5 || k == 7) && (w % 16u) == 0 && aligned16(src) && aligned4(dst) &&
        (!band.top || aligned16(band.top)) && (!band.bottom || aligned16(band.bottom)) && g_variant != 1 &&
        make_conv_round(sum_abs, div, bias, &rnd)) {
        if (k == 3) return conv_fast<3>(rs, dst, w, h, coef, rnd, s);
        if (k == 5) return conv_fast<5>(rs, dst, w, h, coef, rnd, s);
        return conv_fast<7>(rs, dst, w, h, coef, rnd, s);
    }
    ConvCoefGeneric cf;
    for (int i = 0; i < CONV_MAXK * CONV_MAXK; i++) cf.c[i] = i < k * k ? coef[i] : 0;
    dim3 grid((w + CONV_TW - 1) / CONV_TW, (h + CONV_TH - 1) / CONV_TH);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
    size_t smem = (size_t)(CONV_TH + k - 1) * (CONV_TW + k - 1) * 3;
    launch(conv_kernel, dim3(grid), dim3(256), smem, s, rs, dst, w, k, div, bias, cf);
    return PPMX_LAUNCHED();
}

}  // namespace ppmx
