// ppmx_common.cuh -- shared by the kernel translation units: the PDL launch helper, alignment and
// grid helpers, the integer building blocks (exact /3, 4-pixel grey sums, Bayer thresholds) and the
// row resolver for banded rasters.  "ref:N" = /root/reference/ppmx-edward.c line N.
//
// Rasters are flat, packed, row-major: RGB8 = 3 bytes per pixel (the reference's `pixel`,
// ref:39-43), R8 = the .r member only.  Everything integer is done in integers; the two bicubic
// operators run in FP64 with explicit round-to-nearest multiplies and adds (never an FMA) in the
// reference's order.  Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
#pragma once

#include "ppmx_kernels.h"

#include <cuda.h>

#include <atomic>

namespace ppmx {

extern std::atomic<unsigned long long> g_launches;

// SM count of the CURRENT device, cached per device (a racing first call stores the same value twice)
static inline int sm_count()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (!n) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

#define PPMX_LAUNCHED() (cudaGetLastError())

// Programmatic dependent launch (sm_90+): every kernel in this file starts with pdl_wait(), so a
// launch may be scheduled while its predecessor in the stream is still draining; only index
// arithmetic runs before the wait, all global-memory traffic after it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// A load that stays BEHIND pdl_wait().  __ldg / `const __restrict__` loads are ld.global.nc: the compiler treats their target as
// read-only for the kernel's lifetime and may hoist them above the wait (seen with predicated 4-byte __ldg in the tile kernel's
// staging: LDG.E.CONSTANT scheduled before ACQBULK, the predecessor's output read while it was still being written).  A plain
// ld.global in a volatile asm keeps its place; tests/test_sass_checks.py scans every kernel for global accesses before the wait.
__device__ __forceinline__ uint32_t ld_global_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#define PDL_PROLOGUE() \
    do {               \
        pdl_trigger(); \
        pdl_wait();    \
    } while (0)

// opt in to > 48 KB dynamic shared memory once per (kernel, device)
struct SmemOptIn {
    std::atomic<bool> done[64];
};
template <typename K>
static void allow_smem(K kernel, size_t bytes, SmemOptIn &once)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        return;
    }
    if (once.done[dev].load(std::memory_order_acquire)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);  // idempotent: a race repeats it
    once.done[dev].store(true, std::memory_order_release);
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
    at[0].val.programmaticStreamSerializationAllowed = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    g_launches.fetch_add(1, std::memory_order_relaxed);
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

// ceil(matrix * 255) for the Bayer matrix of ref:954 (the products are NOT all integers: 0.125 * 255 = 31.875), in the
// reference's own index order (x%4)*4 + (y%4), ref:967.  Equivalent for an INTEGER grey: grey >= 31.875 <=> grey >= 32,
// so bit = grey < c_bayer  <=>  !(grey >= matrix*255)  (checked for all 256 x 16 cases in tests/test_host_logic.py).
static __constant__ uint8_t c_bayer[16] = {32, 255, 48, 208, 160, 96, 176, 112, 64, 224, 16, 240, 192, 128, 144, 80};

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

// ---- rows of a raster that may be spread over a band and its neighbours' halos ----------------
__host__ __device__ __forceinline__ int mirror_index(int i, int n)
{
    if ((unsigned)i < (unsigned)n) return i;  // inside the raster: the common case costs one compare
    int m = i % (2 * n);
    if (m < 0) m += 2 * n;
    return m < n ? m : 2 * n - 1 - m;
}

// resolves a row of the WHOLE raster to memory: own band, halo above, halo below (maybe peer HBM)
struct RowSource {
    const uint8_t *own, *top, *bottom;
    int y0, h, halo, full_h;
    int lo, hi;  // rows [lo, hi) of the whole raster are readable: the band plus the halo rows that exist
    __device__ __forceinline__ const uint8_t *row(int gy, size_t pitch) const
    {
        gy = mirror_index(gy, full_h);
        return row_plain(gy, pitch);
    }
    // the same without the mirror step, for row numbers that are already inside the raster.  Rows a strip or tile
    // prefetches but never uses (a band whose height is no multiple of the strip height) are clamped to the readable
    // range, so nothing outside [lo, hi) -- possibly a neighbour's HBM -- is ever touched.
    __device__ __forceinline__ const uint8_t *row_plain(int gy, size_t pitch) const
    {
        gy = max(lo, min(gy, hi - 1));
        if (gy >= y0 && gy < y0 + h) return own + (size_t)(gy - y0) * pitch;
        if (gy < y0) return top + (size_t)(gy - (y0 - halo)) * pitch;
        return bottom + (size_t)(gy - (y0 + h)) * pitch;
    }
};

// NOUT words starting `byte_off` (0..15) bytes into the words W[0 .. NOUT + 4]: the word part of the offset is applied as
// two conditional moves per word (by 2 words, then by 1 -- the predicate is the same for the whole warp), the byte part
// as one funnel shift.  (A switch over the four word offsets is compiled to all four cases plus selects: twice the work.)
template <int NIN, int NOUT>
__device__ __forceinline__ void shift_words(const uint32_t (&W)[NIN], uint32_t byte_off, uint32_t (&out)[NOUT])
{
    static_assert(NIN >= NOUT + 4, "shift_words needs four spare words");
    const bool by2 = (byte_off & 8u) != 0, by1 = (byte_off & 4u) != 0;
    const uint32_t sh = 8u * (byte_off & 3u);
    uint32_t X[NOUT + 2], Y[NOUT + 1];
#pragma unroll
    for (int j = 0; j < NOUT + 2; j++) X[j] = by2 ? W[j + 2] : W[j];
#pragma unroll
    for (int j = 0; j < NOUT + 1; j++) Y[j] = by1 ? X[j + 1] : X[j];
#pragma unroll
    for (int k = 0; k < NOUT; k++) out[k] = __funnelshift_r(Y[k], Y[k + 1], sh);
}

// 16 bytes from shared memory at any byte offset `q` of a 16-byte aligned row: two aligned 16-byte reads (consecutive lanes
// read consecutive vectors: no bank conflicts) and a funnel shift; q % 16 is the same for every lane of a run, so the
// word offset is a warp-uniform choice
__device__ __forceinline__ uint4 smem_vec_at(const uint8_t *srow, uint32_t q)
{
    const uint4 *v = reinterpret_cast<const uint4 *>(srow + (q & ~15u));
    const uint4 a = v[0], b = (q & 15u) ? v[1] : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t W[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t o[4];
    shift_words<8, 4>(W, q & 15u, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// `nbytes` bytes from shared memory (16-byte aligned `srow` with 16 bytes of slack behind the run, run at byte `soff`) to
// any global address, by one warp: everything between the first and last 16-byte boundary of the destination as aligned
// vectors, the up to 15 bytes at either end one by one
__device__ __forceinline__ void store_run(uint8_t *g, const uint8_t *srow, uint32_t soff, uint32_t nbytes, uint32_t lane)
{
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);
    const uint32_t hb = min(nbytes, (16u - a) & 15u), nv = (nbytes - hb) >> 4, tb = nbytes - hb - 16u * nv;
    for (uint32_t k = lane; k < nv; k += 32u) reinterpret_cast<uint4 *>(g + hb)[k] = smem_vec_at(srow, soff + hb + 16u * k);
    if (lane < hb + tb) {
        const uint32_t pos = lane < hb ? lane : nbytes - tb + (lane - hb);
        g[pos] = srow[soff + pos];
    }
}

static inline RowSource make_row_source(const uint8_t *src, uint32_t h, const Band &band)
{
    RowSource rs;
    rs.own = src;
    rs.top = band.top;
    rs.bottom = band.bottom;
    rs.y0 = band.full_h ? (int)band.y0 : 0;
    rs.h = (int)h;
    rs.halo = (int)band.halo;
    rs.full_h = band.full_h ? (int)band.full_h : (int)h;
    rs.lo = band.top ? rs.y0 - rs.halo : rs.y0;
    rs.hi = band.bottom ? rs.y0 + rs.h + rs.halo : rs.y0 + rs.h;
    if (rs.lo < 0) rs.lo = 0;
    if (rs.hi > rs.full_h) rs.hi = rs.full_h;
    return rs;
}




}  // namespace ppmx
