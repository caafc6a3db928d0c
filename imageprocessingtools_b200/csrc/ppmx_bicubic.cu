// ppmx_bicubic.cu -- the FP64 operators: bicubic rotate (ref:726-786) and the two imresize passes (ref:820-868), bit-exact.
// Part of libppmx_gpu.so; see ppmx_common.cuh for conventions ("ref:N" = /root/reference/ppmx-edward.c line N).
#include "ppmx_common.cuh"

namespace ppmx {

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
        const bool xu = (CONV == 0) || (CONV == 2 && (i & 1) == 0) || (CONV == 3 && i != 3);  // CONV 3: three of four on the XU pipe
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

// The 4 taps of the bicubic stencil sit at u = floor(n) - 1 + i, so the argument n - u is 1+f, f, f-1, f-2 with
// f = n - floor(n) in [0, 1) (the subtraction is exact): taps 1 and 2 always take the |x| <= 1 polynomial (ref:481-483),
// tap 3 always the 1 < |x| <= 2 one (ref:484-487), and tap 0 too except for f == 0, where |x| == 1 and BOTH polynomials
// give exactly +0.0.  The "r = 0.0 + t" of ref:487 is skipped: t is never -0.0 (a sum that cancels rounds to +0.0).
// Same operations in the same order as `cubic`, minus the branch not taken: 6 resp. 8 FP64 instructions instead of 16.
__device__ __forceinline__ double cubic_near(double x)  // |x| <= 1
{
    const double a1 = fabs(x), a2 = dmul(a1, a1), a3 = dmul(a2, a1);
    return dadd(dsub(dmul(1.5, a3), dmul(2.5, a2)), 1.0);
}
__device__ __forceinline__ double cubic_far(double x)  // 1 <= |x| <= 2
{
    const double a1 = fabs(x), a2 = dmul(a1, a1), a3 = dmul(a2, a1);
    double t = dadd(dmul(-0.5, a3), dmul(2.5, a2));
    t = dsub(t, dmul(4.0, a1));
    return dadd(t, 2.0);
}

// floor of a small double (|v| < 2^31) on the FP64 pipe: v + 1.5*2^52 rounded toward -inf is floor(v) + 1.5*2^52
// exactly; its low word is floor(v) as an int32 and subtracting the constant again gives floor(v) as a double
__device__ __forceinline__ int floor_int(double v) { return __double2loint(__dadd_rd(v, 6755399441055744.0)); }
__device__ __forceinline__ double floor_both(double v, int &i)
{
    const double t = __dadd_rd(v, 6755399441055744.0);
    i = __double2loint(t);
    return dsub(t, 6755399441055744.0);
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
                                                             int xc, int yc, int xo, int yo, double xcd, double ycd)
{
    PDL_PROLOGUE();
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nw || y >= nh) return;
    uint8_t *out = dst + ((size_t)y * nw + x) * 3;

    const int x0 = ((int)x - xo) - xc, y0 = ((int)y - yo) - yc;                              // ref:731-735
    const double xd = (double)x0, yd = (double)y0;
    const double nX = dadd(dadd(dmul(cs, xd), dmul(sn, yd)), xcd);     // ref:741
    const double nY = dadd(dadd(-dmul(sn, xd), dmul(cs, yd)), ycd);    // ref:742
    uint32_t r = 0, g = 0, b = 0;  // uncovered output stays 0 (ref:727)
    // round(n) = floor(n + 0.5) (ref:27) is integer-valued and far inside the int range, so it is taken as the low word
    // of (n + 0.5) + 1.5*2^52 rounded down -- no FRND/F2I on the 16-lane XU pipe -- and the comparisons of ref:744,752
    // are done on integers (the alu pipe is idle, the FP64 pipe is the bottleneck); w - 2u wraps for w < 2 exactly
    // like the reference's unsigned arithmetic
    const int irx = floor_int(dadd(nX, 0.5)), iry = floor_int(dadd(nY, 0.5));
    if ((uint32_t)irx < w && (uint32_t)iry < h) {                                              // ref:744
        if (irx > 1 && iry > 1 && (uint32_t)irx < w - 2u && (uint32_t)iry < h - 2u) {          // ref:752
            int fxi, fyi;
            const double fx = floor_both(nX, fxi), fy = floor_both(nY, fyi);
            double wx[4], wy[4];
            const int u0 = fxi - 1, v0 = fyi - 1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                // ref:758,761 compute u = floor(n) - 1 + i in double, cast it to int and convert it back for n - u:
                // the double is integer-valued, so the round trip (two XU-pipe conversions) is the identity
                const double u = dadd(dsub(fx, 1.0), (double)i), v = dadd(dsub(fy, 1.0), (double)i);
                wx[i] = (i == 1 || i == 2) ? cubic_near(dsub(nX, u)) : cubic_far(dsub(nX, u));
                wy[i] = (i == 1 || i == 2) ? cubic_near(dsub(nY, v)) : cubic_far(dsub(nY, v));
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
                    // (0.0 + x == x up to the sign of a zero, which no later step can see: first terms are not added)
                    p0 = dmul(d[0][0], wx[0]);
                    p1 = dmul(d[0][1], wx[0]);
                    p2 = dmul(d[0][2], wx[0]);
#pragma unroll
                    for (int i = 1; i < 4; i++) {  // ref:762-764
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
                q0 = j ? dadd(q0, dmul(p0, wy[j])) : dmul(p0, wy[0]);  // ref:766-768
                q1 = j ? dadd(q1, dmul(p1, wy[j])) : dmul(p1, wy[0]);
                q2 = j ? dadd(q2, dmul(p2, wy[j])) : dmul(p2, wy[0]);
            }
            // ref:771-781: q < 0 -> 0, q >= 256 -> 255, else truncate.  floor(q) = the low word of q + 1.5*2^52 rounded
            // down (|q| is small); floor(q) < 0 <=> q < 0 and floor(q) >= 256 <=> q >= 256, so the clamp is an integer one
            r = (uint32_t)min(max(__double2loint(__dadd_rd(q0, 6755399441055744.0)), 0), 255);
            g = (uint32_t)min(max(__double2loint(__dadd_rd(q1, 6755399441055744.0)), 0), 255);
            b = (uint32_t)min(max(__double2loint(__dadd_rd(q2, 6755399441055744.0)), 0), 255);
        } else {  // nearest, ref:783
            const uint8_t *p = src + ((size_t)iry * w + (size_t)irx) * 3;
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
    if ((w % 4u) == 0 && aligned4(src) && PPMX_VARIANT != 1) {
        // with the index conversions off the XU pipe (floor_int/floor_both) all 48 byte conversions fit there:
        // measured 56.5 Gpix/s (CONV 0) vs 53.5 (3 of 4 on XU) vs 51.0 (half) vs 47.8 (all on the FP64 pipe)
#ifdef PPMX_TUNING  // other splits of the byte conversions between the XU and FP64 pipes (profiles/r1_sweep_fp64.txt)
        if (PPMX_VARIANT == 2)
            launch(rotate_bicubic_kernel<true, 1>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo, (double)xc, (double)yc);
        else if (PPMX_VARIANT == 3)
            launch(rotate_bicubic_kernel<true, 2>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo, (double)xc, (double)yc);
        else if (PPMX_VARIANT == 4)
            launch(rotate_bicubic_kernel<true, 3>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo, (double)xc, (double)yc);
        else
#endif
            launch(rotate_bicubic_kernel<true, 0>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo, (double)xc, (double)yc);
    } else {
        launch(rotate_bicubic_kernel<false, 1>, grid, block, 0, s, src, dst, w, h, nw, nh, cos_t, sin_t, xc, yc, xo, yo, (double)xc, (double)yc);
    }
    return PPMX_LAUNCHED();
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

// four results at once: floor(s + 0.5) as above, then ONE saturating pack per two values (I2IP) does the
// "< 0 -> 0, >= 256 -> 255" clamp of ref:835 and the byte packing together
__device__ __forceinline__ int quantise_floor(double s)
{
    return __double2loint(__dadd_rd(dadd(s, 0.5), 6755399441055744.0));
}
__device__ __forceinline__ uint32_t quantise_pack4(double s0, double s1, double s2, double s3)
{
    uint32_t hi, out;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(quantise_floor(s3)), "r"(quantise_floor(s2)), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(quantise_floor(s1)), "r"(quantise_floor(s0)), "r"(hi));
    return out;
}

// height pass, fast path (row pitch % 16 == 0, aligned, K <= 64): one thread = 16 bytes of an
// output row.  The row's K weights and source-row numbers are staged once per CTA in shared memory
// so the K source loads of a thread are independent of each other and fly four at a time.
constexpr int ROWS16_MAXK = 64;
template <int NW> struct RowVec;
template <> struct RowVec<4> { typedef uint4 type; };
template <> struct RowVec<2> { typedef uint2 type; };
__device__ __forceinline__ void vec_words(const uint4 &v, uint32_t (&w)[4]) { w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
__device__ __forceinline__ void vec_words(const uint2 &v, uint32_t (&w)[2]) { w[0] = v.x; w[1] = v.y; }
__device__ __forceinline__ void words_vec(const uint32_t (&w)[4], uint4 &v) { v = make_uint4(w[0], w[1], w[2], w[3]); }
__device__ __forceinline__ void words_vec(const uint32_t (&w)[2], uint2 &v) { v = make_uint2(w[0], w[1]); }

// NW = words per thread: 4 (16 bytes, 57-64 registers) or 2 (8 bytes, fewer registers, more warps in flight)
// KT > 0: the tap count as a compile-time constant (4 for every upscale, up to 8 down to x0.5): the group loop and
// its bounds tests fold away (measured +8.5 % at K = 4)
template <int CONV, int NW, int KT = 0>
__global__ void __launch_bounds__(256) imresize_rows16_kernel(const RowSource src, uint8_t *__restrict__ dst,
                                                              uint32_t row_vecs, int taps,
                                                              const double *__restrict__ wts, const int *__restrict__ idx)
{
    typedef typename RowVec<NW>::type vec_t;
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
    const size_t row_bytes = (size_t)row_vecs * (4 * NW);

    double acc[4 * NW];
#pragma unroll
    for (int i = 0; i < 4 * NW; i++) acc[i] = 0.0;
    const int ntaps = KT ? KT : taps;
    auto group = [&](int z0) {  // four taps: their source vectors fly together
        vec_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (z0 + u < ntaps) v[u] = __ldg(reinterpret_cast<const vec_t *>(src.row_plain(s_i[z0 + u], row_bytes)) + xv);
#pragma unroll
        for (int u = 0; u < 4; u++) {  // tap order is the reference's summation order (ref:826-830)
            if (z0 + u >= ntaps) break;
            const double wz = s_w[z0 + u];
            uint32_t wd[NW];
            vec_words(v[u], wd);
#pragma unroll
            for (int q = 0; q < NW; q++) {
                double d[4];
                word_to_double4<CONV>(wd[q], d);
#pragma unroll
                for (int b = 0; b < 4; b++)  // the first tap is not added to 0.0 (same value; a zero's sign is never seen)
                    acc[4 * q + b] = (z0 + u) ? dadd(acc[4 * q + b], dmul(d[b], wz)) : dmul(d[b], wz);
            }
        }
    };
    if (KT) {
#pragma unroll
        for (int z0 = 0; z0 < KT; z0 += 4) group(z0);
    } else {
#pragma unroll 1
        for (int z0 = 0; z0 < ntaps; z0 += 4) group(z0);
    }
    uint32_t o[NW];
#pragma unroll
    for (int q = 0; q < NW; q++)
        o[q] = quantise_pack4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    vec_t ov;
    words_vec(o, ov);
    reinterpret_cast<vec_t *>(dst + (size_t)y * row_bytes)[xv] = ov;
}

// Height pass, TWO output rows per CTA (a thread: 8 bytes of each).  Away from the raster's top and bottom the K taps of an
// output row are K consecutive source rows, and the next output row's taps start `delta` rows further (0 or 1 when
// upscaling, about 1/scale when downscaling): the pair reads K + delta distinct source rows instead of 2 K, and every source
// byte is converted to double ONCE for both rows -- the conversions (XU pipe, 16 lanes/clk/SM) are what kept the FP64 pipe at
// 55 % in the one-row kernel.  Each row's sum still runs over its own taps in tap order with separately rounded products
// and sums (ref:826-830), so the bytes are the same.  Pairs whose taps are not consecutive rows (mirrored taps at the
// raster's ends) and a last odd row take the one-row arithmetic inside the same kernel.
// OWN: all K + DELTA rows lie in the band's own raster (always, unless a neighbour's halo rows are involved): plain
// pointer steps instead of the band resolver
template <int CONV, int KT, int DELTA, bool OWN>
__device__ __forceinline__ void rows_pair(const RowSource &src, int base, size_t row_bytes, uint32_t xv, const double *wa,
                                          const double *wb, double (&accA)[8], double (&accB)[8])
{
    constexpr int U = KT + DELTA;  // distinct source rows of the pair
    const uint8_t *p0 = OWN ? src.own + (size_t)(base - src.y0) * row_bytes + (size_t)xv * 8 : nullptr;
#pragma unroll
    for (int u0 = 0; u0 < U; u0 += 4) {
        uint2 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (u0 + q < U)
                v[q] = OWN ? __ldg(reinterpret_cast<const uint2 *>(p0 + (size_t)(u0 + q) * row_bytes))
                           : __ldg(reinterpret_cast<const uint2 *>(src.row_plain(base + u0 + q, row_bytes)) + xv);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int u = u0 + q;
            if (u >= U) break;
            double d[8];
            {
                double t[4];
                word_to_double4<CONV>(v[q].x, t);
                d[0] = t[0], d[1] = t[1], d[2] = t[2], d[3] = t[3];
                word_to_double4<CONV>(v[q].y, t);
                d[4] = t[0], d[5] = t[1], d[6] = t[2], d[7] = t[3];
            }
            if (u < KT) {  // tap u of the first row (tap order = source-row order, ref:826-830)
                const double w = wa[u];
#pragma unroll
                for (int b = 0; b < 8; b++) accA[b] = u ? dadd(accA[b], dmul(d[b], w)) : dmul(d[b], w);
            }
            if (u >= DELTA) {  // tap u - DELTA of the second row
                const double w = wb[u - DELTA];
#pragma unroll
                for (int b = 0; b < 8; b++) accB[b] = (u > DELTA) ? dadd(accB[b], dmul(d[b], w)) : dmul(d[b], w);
            }
        }
    }
}

template <int CONV, int KT>
__global__ void __launch_bounds__(256) imresize_rows8x2_kernel(const RowSource src, uint8_t *__restrict__ dst, uint32_t row_vecs,
                                                               int out_rows, const double *__restrict__ wts,
                                                               const int *__restrict__ idx)
{
    PDL_PROLOGUE();
    __shared__ double s_w[2][KT];
    __shared__ int s_i[2][KT];
    const int y0 = 2 * blockIdx.y;
    const bool two = y0 + 1 < out_rows;
    if ((int)threadIdx.x < 2 * KT) {
        const int r = threadIdx.x / KT, z = threadIdx.x % KT;
        if (r == 0 || two) {
            s_w[r][z] = __ldg(wts + (size_t)(y0 + r) * KT + z);
            s_i[r][z] = __ldg(idx + (size_t)(y0 + r) * KT + z);
        }
    }
    __syncthreads();
    const uint32_t xv = blockIdx.x * 256 + threadIdx.x;
    if (xv >= row_vecs) return;
    const size_t row_bytes = (size_t)row_vecs * 8;
    // shared path: both rows' taps are consecutive source rows and the second row's start `delta` rows after the first's
    const int base = s_i[0][0];
    int delta = two ? s_i[1][0] - base : -1;
    bool shared = two && delta >= 0 && delta <= KT;
#pragma unroll
    for (int z = 1; z < KT; z++) shared = shared && s_i[0][z] == base + z && (!two || s_i[1][z] == s_i[1][0] + z);

    double accA[8], accB[8];
#pragma unroll
    for (int b = 0; b < 8; b++) accB[b] = 0.0;  // (x + 0.0 == x up to the sign of a zero, which the rounding never sees)
    if (shared && delta <= 3) {
        // the pair's K + delta source rows with delta a COMPILE-TIME constant per case: no branch inside, every load and
        // conversion of the pair can be in flight together
        const bool own = base >= src.y0 && base + KT + delta <= src.y0 + src.h;
#define PPMX_PAIR(D)                                                                                   \
    if (own) rows_pair<CONV, KT, D, true>(src, base, row_bytes, xv, s_w[0], s_w[1], accA, accB);        \
    else rows_pair<CONV, KT, D, false>(src, base, row_bytes, xv, s_w[0], s_w[1], accA, accB)
        switch (delta) {
        case 0: PPMX_PAIR(0); break;
        case 1: PPMX_PAIR(1); break;
        case 2: PPMX_PAIR(2); break;
        default: PPMX_PAIR(3); break;
        }
#undef PPMX_PAIR
    } else {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            if (r == 1 && !two) break;
            double(&acc)[8] = r ? accB : accA;
#pragma unroll
            for (int z0 = 0; z0 < KT; z0 += 4) {
                uint2 v[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (z0 + q < KT) v[q] = __ldg(reinterpret_cast<const uint2 *>(src.row_plain(s_i[r][z0 + q], row_bytes)) + xv);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int z = z0 + q;
                    if (z >= KT) break;
                    const double wz = s_w[r][z];
                    double t[4];
                    word_to_double4<CONV>(v[q].x, t);
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[b] = z ? dadd(acc[b], dmul(t[b], wz)) : dmul(t[b], wz);
                    word_to_double4<CONV>(v[q].y, t);
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[4 + b] = z ? dadd(acc[4 + b], dmul(t[b], wz)) : dmul(t[b], wz);
                }
            }
        }
    }
    uint2 o;
    o.x = quantise_pack4(accA[0], accA[1], accA[2], accA[3]);
    o.y = quantise_pack4(accA[4], accA[5], accA[6], accA[7]);
    reinterpret_cast<uint2 *>(dst + (size_t)y0 * row_bytes)[xv] = o;
    if (two) {
        o.x = quantise_pack4(accB[0], accB[1], accB[2], accB[3]);
        o.y = quantise_pack4(accB[4], accB[5], accB[6], accB[7]);
        reinterpret_cast<uint2 *>(dst + (size_t)(y0 + 1) * row_bytes)[xv] = o;
    }
}

template <int KT>
static void launch_rows8x2(const RowSource &src, uint8_t *dst, uint32_t row_bytes, int rows, const double *w0, const int *i0,
                           cudaStream_t s)
{
    const uint32_t vecs = row_bytes / 8;
    for (int y0 = 0; y0 < rows; y0 += 2 * 65535) {
        const int n = min(2 * 65535, rows - y0);
        dim3 grid((vecs + 255) / 256, (n + 1) / 2);
        launch(imresize_rows8x2_kernel<3, KT>, grid, dim3(256), 0, s, src, dst + (size_t)y0 * row_bytes, vecs, n, w0 + (size_t)y0 * KT,
               i0 + (size_t)y0 * KT);
    }
}

// (A pipelined variant -- a CTA owning 8 consecutive output rows, tables staged once, the next group of four source
// vectors in flight in a second register buffer while the current one is multiplied -- measured SLOWER than the
// one-row-per-CTA kernel above: 67.6 vs 73.1 Gpix/s at x1.5 and 135 vs 178 at x0.5, profiles/r1_sweep_fp64.txt.
// Many short CTAs hide the two trips to memory better than a software pipeline at 78 registers.)

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
        // the words of row y+1 are requested before row y is evaluated; two register sets alternate (no copies)
        const uint32_t *rw0 = reinterpret_cast<const uint32_t *>(src) + w0;
        const size_t pitch_words = in_pitch / 4;
        auto fetch = [&](uint32_t y, uint32_t(&q)[NS + 1]) {
            const uint32_t *rw = rw0 + (size_t)y * pitch_words;
#pragma unroll
            for (int j = 0; j <= NS; j++)  // word j is needed iff it starts before the last tap byte
                q[j] = (y < y1 && 4u * j < (b0 & 3u) + 3u * K) ? __ldg(rw + j) : 0u;
        };
        auto eval = [&](uint32_t y, const uint32_t(&q)[NS + 1]) {
            double d[NS][4];
#pragma unroll
            for (int j = 0; j < NS; j++) word_to_double4<CONV>(__funnelshift_r(q[j], q[j + 1], sh), d[j]);
            double s0 = dmul(d[0][0], wk[0]), s1 = dmul(d[0][1], wk[0]), s2 = dmul(d[0][2], wk[0]);  // 0.0 + x == x
#pragma unroll
            for (int z = 1; z < K; z++) {  // ref:852-858, tap order
                s0 = dadd(s0, dmul(d[(3 * z) >> 2][(3 * z) & 3], wk[z]));
                s1 = dadd(s1, dmul(d[(3 * z + 1) >> 2][(3 * z + 1) & 3], wk[z]));
                s2 = dadd(s2, dmul(d[(3 * z + 2) >> 2][(3 * z + 2) & 3], wk[z]));
            }
            uint8_t *o = dst + (size_t)y * out_pitch + (size_t)x * 3;
            o[0] = (uint8_t)quantise_fast(s0);
            o[1] = (uint8_t)quantise_fast(s1);
            o[2] = (uint8_t)quantise_fast(s2);
        };
        // three register sets rotate: the words of rows y + 1 and y + 2 are in flight while row y is evaluated (with two
        // sets the kernel waited on memory: long-scoreboard stalls 5.1 per issue, profiles/r2_ncu_*.txt)
        uint32_t qa[NS + 1], qb[NS + 1], qc[NS + 1];
        fetch(y0, qa);
        fetch(y0 + 1, qb);
        for (uint32_t y = y0; y < y1; y += 3) {
            fetch(y + 2, qc);
            eval(y, qa);
            if (y + 1 >= y1) break;
            fetch(y + 3, qa);
            eval(y + 1, qb);
            if (y + 2 >= y1) break;
            fetch(y + 4, qb);
            eval(y + 2, qc);
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
#ifdef PPMX_TUNING
        if (PPMX_VARIANT == 2)
            launch(imresize_colsK_kernel<K, 1>, grid, dim3(128), 0, s, src + (size_t)y0 * w * 3,
                   dst + (size_t)y0 * out_w * 3, w, rows, out_w, rows_per_cta, wts, idx);
        else if (PPMX_VARIANT == 3)
            launch(imresize_colsK_kernel<K, 0>, grid, dim3(128), 0, s, src + (size_t)y0 * w * 3,
                   dst + (size_t)y0 * out_w * 3, w, rows, out_w, rows_per_cta, wts, idx);
        else if (PPMX_VARIANT == 4)
            launch(imresize_colsK_kernel<K, 2>, grid, dim3(128), 0, s, src + (size_t)y0 * w * 3,
                   dst + (size_t)y0 * out_w * 3, w, rows, out_w, rows_per_cta, wts, idx);
        else  // three of four byte conversions on the XU pipe: 1-3 % faster than half/half once the first-tap adds are gone
#endif
            launch(imresize_colsK_kernel<K, 3>, grid, dim3(128), 0, s, src + (size_t)y0 * w * 3,
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
        if (row_bytes % 16 == 0 && aligned16(src_ptr) && aligned16(dst) && halo_ok && taps <= ROWS16_MAXK && PPMX_VARIANT != 1) {
#ifdef PPMX_TUNING
            // two output rows per CTA sharing their source rows' conversions: measured 15 % SLOWER than one row per CTA
            // (x1.5: 67.3 vs 79.0 Gpix/s over both passes, profiles/r2_sweep_fp64.txt) -- kept for the record only
            if (taps >= 4 && taps <= 8 && PPMX_VARIANT == 9) {
                switch (taps) {
                case 4: launch_rows8x2<4>(src, dst, row_bytes, out_size, d_weights, d_indices, s); break;
                case 5: launch_rows8x2<5>(src, dst, row_bytes, out_size, d_weights, d_indices, s); break;
                case 6: launch_rows8x2<6>(src, dst, row_bytes, out_size, d_weights, d_indices, s); break;
                case 7: launch_rows8x2<7>(src, dst, row_bytes, out_size, d_weights, d_indices, s); break;
                default: launch_rows8x2<8>(src, dst, row_bytes, out_size, d_weights, d_indices, s); break;
                }
                return cudaGetLastError();
            }
#endif
            [[maybe_unused]] const bool narrow = (PPMX_VARIANT == 6);  // 8 bytes per thread
            const uint32_t vecs = narrow ? row_bytes / 8 : row_bytes / 16;
            dim3 grid((vecs + 255) / 256, 1);
            for (int y0 = 0; y0 < out_size; y0 += 65535) {
                int rows = min(65535, out_size - y0);
                grid.y = rows;
                uint8_t *d0 = dst + (size_t)y0 * row_bytes;
                const double *w0 = d_weights + (size_t)y0 * taps;
                const int *i0 = d_indices + (size_t)y0 * taps;
#ifdef PPMX_TUNING
                if (narrow) launch(imresize_rows16_kernel<2, 2>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else if (PPMX_VARIANT == 2) launch(imresize_rows16_kernel<1, 4>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else if (PPMX_VARIANT == 3) launch(imresize_rows16_kernel<0, 4>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else if (PPMX_VARIANT == 4) launch(imresize_rows16_kernel<2, 4>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else
#endif
                if (taps == 4 && PPMX_VARIANT != 5) launch(imresize_rows16_kernel<3, 4, 4>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else if (taps == 5 && PPMX_VARIANT != 5) launch(imresize_rows16_kernel<3, 4, 5>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else if (taps == 6 && PPMX_VARIANT != 5) launch(imresize_rows16_kernel<3, 4, 6>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else if (taps == 7 && PPMX_VARIANT != 5) launch(imresize_rows16_kernel<3, 4, 7>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else if (taps == 8 && PPMX_VARIANT != 5) launch(imresize_rows16_kernel<3, 4, 8>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
                else launch(imresize_rows16_kernel<3, 4>, grid, dim3(256), 0, s, src, d0, vecs, taps, w0, i0);
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
    if ((w % 4u) == 0 && aligned4(src) && taps >= 4 && taps <= 8 && PPMX_VARIANT != 1) {
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


}  // namespace ppmx
