// ppmx_color.cu -- gray (ref:998-1000), mono (ref:964-969), the P4 packer (ref:268-284), .r extraction (ref:263-267) and the fused histogram (extension).
// Part of libppmx_gpu.so; see ppmx_common.cuh for conventions ("ref:N" = /root/reference/ppmx-edward.c line N).
#include "ppmx_common.cuh"

namespace ppmx {

#ifdef PPMX_TUNING
int g_variant = 0;  // ppmx_gpu_set_tuning("variant", n): only in the tuning build (libppmx_gpu_tuning.so)
#endif
std::atomic<int> g_pdl{1};  // ppmx_gpu_set_tuning("pdl", 0/1); process-wide
std::atomic<unsigned long long> g_launches{0};
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void add_launches(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

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
[[maybe_unused]] constexpr int HP_MAX_GROUPS = 15;
[[maybe_unused]] constexpr size_t HP_SMEM = (64 * HP_THREADS + 256) * sizeof(uint32_t);

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

#ifdef PPMX_TUNING
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

#endif  // PPMX_TUNING

template <bool HIST, bool STORE>
static cudaError_t gray_dispatch(const uint8_t *src, uint8_t *dst, size_t npix, unsigned long long *d_hist,
                                 cudaStream_t s)
{
    if (npix == 0) return cudaSuccess;
    if (aligned16(src) && (!STORE || aligned16(dst))) {
        size_t ngroups = npix / 16;
#ifdef PPMX_TUNING  // alternative implementations, measured and not chosen (profiles/r1_sweep_variants.txt): tuning build only
        if (HIST && PPMX_VARIANT == 5) {
            // thread-private byte counters: <= 15 groups per thread between folds, 3 CTAs per SM
            static SmemOptIn ok;
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
        if (!HIST && (PPMX_VARIANT == 6 || PPMX_VARIANT == 7)) {  // smaller CTAs: shorter tail, more CTA launches
            const unsigned blk = PPMX_VARIANT == 6 ? 128u : 64u;
            unsigned grid = (unsigned)((ngroups + blk - 1) / blk);
            if (blk == 128) launch(gray_flat_kernel<128>, dim3(grid ? grid : 1), dim3(128), 0, s,
                                   reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), ngroups, npix);
            else launch(gray_flat_kernel<64>, dim3(grid ? grid : 1), dim3(64), 0, s,
                        reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), ngroups, npix);
            return PPMX_LAUNCHED();
        }
        if (!HIST && PPMX_VARIANT == 4) {  // cp.async.bulk + mbarrier ring
            static SmemOptIn ok;
            allow_smem(gray_tma_kernel, GT_SMEM, ok);
            size_t ntiles = (ngroups + GT_TILE_GROUPS - 1) / GT_TILE_GROUPS;
            unsigned grid = (unsigned)(ntiles < (size_t)sm_count() * 4 ? (ntiles ? ntiles : 1) : (size_t)sm_count() * 4);
            launch(gray_tma_kernel, dim3(grid), dim3(256), GT_SMEM, s, src, reinterpret_cast<uint4 *>(dst), ngroups, npix);
            return PPMX_LAUNCHED();
        }
        if ((HIST && PPMX_VARIANT == 2) || (!HIST && PPMX_VARIANT == 1)) {  // persistent grid-stride grid, per-warp bins
            unsigned grid = wave_grid(ngroups ? ngroups : 1, 256, 8);
            launch(gray_vec_kernel<HIST, STORE>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const uint4 *>(src),
                   reinterpret_cast<uint4 *>(dst), ngroups, npix, d_hist);
            return PPMX_LAUNCHED();
        }
#endif
        if (HIST) {
            // lane-private columns, RED.shared; one 1024-thread CTA per SM
            static SmemOptIn ok;
            allow_smem(gray_hist_lanes_kernel<STORE>, HL_SMEM, ok);
            size_t want = (ngroups + HL_THREADS - 1) / HL_THREADS, wave = (size_t)sm_count();
            unsigned grid = (unsigned)(want < 1 ? 1 : want < wave ? want : wave);
            launch(gray_hist_lanes_kernel<STORE>, dim3(grid), dim3(HL_THREADS), HL_SMEM, s, reinterpret_cast<const uint4 *>(src),
                   reinterpret_cast<uint4 *>(dst), ngroups, npix, d_hist);
            return PPMX_LAUNCHED();
        }
        // one 16-pixel group per thread, one CTA per 256 groups, no loop
        unsigned grid = (unsigned)((ngroups + 255) / 256);
        launch(gray_flat_kernel<256>, dim3(grid ? grid : 1), dim3(256), 0, s, reinterpret_cast<const uint4 *>(src),
               reinterpret_cast<uint4 *>(dst), ngroups, npix);
    } else if (!HIST && STORE && PPMX_VARIANT == 0 && npix >= 8192) {
        // odd pointers (e.g. raster b of a batch whose size is no multiple of 16): a flat operator may view the pixel run as
        // rows of any length -- 4096-pixel rows through the tile kernel of ppmx_fused.cu, the remainder one byte per thread
        const uint32_t vw = 4096;
        const size_t vh = npix / vw, rest = npix - vh * vw;
        GeomOp go = {};
        go.point = 1;
        for (size_t y = 0; y < vh; y += 65535u * 64u) {
            const uint32_t rows = (uint32_t)(vh - y < 65535u * 64u ? vh - y : 65535u * 64u);
            cudaError_t e = geom_point(src + y * vw * 3, dst + y * vw, vw, rows, 0, go, s);
            if (e != cudaSuccess) return e;
        }
        if (rest)
            launch(gray_scalar_kernel<false, true>, dim3(wave_grid(rest, 256, 8)), dim3(256), 0, s, src + vh * vw * 3, dst + vh * vw, rest,
                   (unsigned long long *)nullptr);
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
    } else if (PPMX_VARIANT == 0 && (size_t)w * h >= 4096 && h <= 65535u * 64u) {
        // any width / alignment: the tile kernel of ppmx_fused.cu (rows in as aligned words, bits out as shifted words)
        GeomOp go = {};
        go.point = 3;
        go.my_add = (int)(y0 & 3u);
        return geom_point(src, dst, w, h, 0, go, s);
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
    if (PPMX_VARIANT == 0 && npix >= 8192) {  // the writer's .r gather as rows of 4096 pixels through the tile kernel
        const uint32_t vw = 4096;
        const size_t vh = npix / vw, rest = npix - vh * vw;
        GeomOp go = {};
        go.point = 2;
        for (size_t y = 0; y < vh; y += 65535u * 64u) {
            const uint32_t rows = (uint32_t)(vh - y < 65535u * 64u ? vh - y : 65535u * 64u);
            cudaError_t e = geom_point(src + y * vw * 3, dst + y * vw, vw, rows, 0, go, s);
            if (e != cudaSuccess) return e;
        }
        if (rest) launch(extract_r_kernel, dim3(wave_grid(rest, 256, 8)), dim3(256), 0, s, src + vh * vw * 3, dst + vh * vw, rest);
        return PPMX_LAUNCHED();
    }
    launch(extract_r_kernel, dim3(wave_grid(npix, 256, 8)), dim3(256), 0, s, src, dst, npix);
    return PPMX_LAUNCHED();
}



// ------------------------------------------------------------------------------------------
// EXTENSION (no reference counterpart, parity unpinned): levels = every byte through a 256-entry table.
// The table travels as a kernel argument (256 B) and is unpacked to one word per entry in shared memory (1 KB), so
// entry v sits in bank v % 32 and two lanes collide only when their bytes differ by a multiple of 32.  A thread maps
// one 16-byte vector: 16 LDS, 12 PRMT to pack, HBM-bound (6 B/px on RGB8).
// ------------------------------------------------------------------------------------------
struct LevelsLut {
    uint32_t w[64];  // 256 entries, 4 per word
};

__device__ __forceinline__ uint32_t levels_word(const uint32_t *lut_s, uint32_t v)
{
    const uint32_t a = lut_s[v & 0xFFu], b = lut_s[(v >> 8) & 0xFFu], c = lut_s[(v >> 16) & 0xFFu], d = lut_s[v >> 24];
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

__global__ void __launch_bounds__(256) levels_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t nbytes,
                                                     const LevelsLut lut)
{
    pdl_trigger();
    __shared__ uint32_t lut_s[256];
    lut_s[threadIdx.x] = (lut.w[threadIdx.x >> 2] >> (8 * (threadIdx.x & 3))) & 0xFFu;
    __syncthreads();
    const size_t nvec = nbytes / 16, g = (size_t)blockIdx.x * 256 + threadIdx.x;
    pdl_wait();
    if (g < nvec) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + g);
        uint4 o;
        o.x = levels_word(lut_s, v.x);
        o.y = levels_word(lut_s, v.y);
        o.z = levels_word(lut_s, v.z);
        o.w = levels_word(lut_s, v.w);
        reinterpret_cast<uint4 *>(dst)[g] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < nbytes - nvec * 16) {  // ragged tail
        const size_t i = nvec * 16 + threadIdx.x;
        dst[i] = (uint8_t)lut_s[src[i]];
    }
}

// The same with one table COLUMN PER LANE, like the histogram: lut_l[v][lane] (32 KB), so lane l only reads bank l
// and the 16 lookups of a vector never collide, whatever the pixel values; one fat CTA per SM fills its copy once
// and strides over the raster with two vectors in flight per thread.  (The one-word-per-entry kernel above
// serialises ~3.5 ways on noise: 0.52 of the HBM roofline; it remains for small rasters and as variant 1.)
constexpr int LV_THREADS = 1024;
constexpr size_t LV_SMEM = 256 * 32 * sizeof(uint32_t);

__device__ __forceinline__ uint32_t levels_word_lane(const uint32_t *col, uint32_t v)
{
    const uint32_t a = col[(v & 0xFFu) << 5], b = col[((v >> 8) & 0xFFu) << 5], c = col[((v >> 16) & 0xFFu) << 5], d = col[(v >> 24) << 5];
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
__device__ __forceinline__ uint4 levels_vec_lane(const uint32_t *col, uint4 v)
{
    return make_uint4(levels_word_lane(col, v.x), levels_word_lane(col, v.y), levels_word_lane(col, v.z), levels_word_lane(col, v.w));
}

__global__ void __launch_bounds__(LV_THREADS, 1) levels_lanes_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                                     size_t nvec, size_t nbytes, const LevelsLut lut)
{
    pdl_trigger();
    extern __shared__ __align__(16) uint32_t lut_l[];
    for (uint32_t i = threadIdx.x; i < 256 * 32; i += LV_THREADS) {
        const uint32_t v = i >> 5;
        lut_l[i] = (lut.w[v >> 2] >> (8 * (v & 3))) & 0xFFu;
    }
    __syncthreads();
    const uint32_t *col = lut_l + (threadIdx.x & 31u);
    pdl_wait();
    const size_t stride = (size_t)gridDim.x * LV_THREADS;
    size_t g = (size_t)blockIdx.x * LV_THREADS + threadIdx.x;
    for (; g + stride < nvec; g += 2 * stride) {
        const uint4 a = __ldg(src + g), b = __ldg(src + g + stride);
        dst[g] = levels_vec_lane(col, a);
        dst[g + stride] = levels_vec_lane(col, b);
    }
    if (g < nvec) dst[g] = levels_vec_lane(col, __ldg(src + g));
    if (blockIdx.x == 0 && threadIdx.x < nbytes - nvec * 16) {  // ragged tail
        const size_t i = nvec * 16 + threadIdx.x;
        reinterpret_cast<uint8_t *>(dst)[i] = (uint8_t)col[(uint32_t)reinterpret_cast<const uint8_t *>(src)[i] << 5];
    }
}

__global__ void __launch_bounds__(256) levels_bytes_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                           size_t nbytes, const LevelsLut lut)
{  // unaligned rasters
    PDL_PROLOGUE();
    __shared__ uint32_t lut_s[256];
    lut_s[threadIdx.x] = (lut.w[threadIdx.x >> 2] >> (8 * (threadIdx.x & 3))) & 0xFFu;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nbytes; i += (size_t)gridDim.x * 256) dst[i] = (uint8_t)lut_s[src[i]];
}

cudaError_t levels(const uint8_t *src, uint8_t *dst, size_t nbytes, const uint8_t *lut, cudaStream_t s)
{
    if (!lut) return cudaErrorInvalidValue;
    if (!nbytes) return cudaSuccess;
    LevelsLut l;
    for (int i = 0; i < 64; i++)
        l.w[i] = (uint32_t)lut[4 * i] | (uint32_t)lut[4 * i + 1] << 8 | (uint32_t)lut[4 * i + 2] << 16 | (uint32_t)lut[4 * i + 3] << 24;
    if (aligned16(src) && aligned16(dst) && nbytes >= (size_t)1 << 20 && PPMX_VARIANT != 1) {  // (variant: tuning build only)
        static SmemOptIn ok;
        allow_smem(levels_lanes_kernel, LV_SMEM, ok);
        const size_t nvec = nbytes / 16, want = (nvec + LV_THREADS - 1) / LV_THREADS, wave = (size_t)sm_count();
        launch(levels_lanes_kernel, dim3((unsigned)(want < wave ? want : wave)), dim3(LV_THREADS), LV_SMEM, s,
               reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), nvec, nbytes, l);
    } else if (aligned16(src) && aligned16(dst) && nbytes >= 16) {
        const size_t nvec = nbytes / 16, ctas = (nvec + 255) / 256;
        if (ctas > 0x7FFFFFFFull) return cudaErrorInvalidValue;
        launch(levels_kernel, dim3((unsigned)ctas), dim3(256), 0, s, src, dst, nbytes, l);
    } else {
        launch(levels_bytes_kernel, dim3(wave_grid(nbytes, 256, 8)), dim3(256), 0, s, src, dst, nbytes, l);
    }
    return PPMX_LAUNCHED();
}

}  // namespace ppmx
