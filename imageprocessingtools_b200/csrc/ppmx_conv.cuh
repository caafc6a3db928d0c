// ppmx_conv.cuh -- shared by the convolution translation units (ppmx_conv.cu, ppmx_conv_sep.cu): rounding of the
// quotient, the mixed-sign dp4a, the rank-1 factorisation.  EXTENSION operators (no reference counterpart).
#pragma once
#include "ppmx_common.cuh"

namespace ppmx {

constexpr int CONV_MAXK = 15;

// exact floor((2*acc + div) / (2*div)) + bias with one multiply-high: the numerator is shifted to
// be non-negative by a multiple K of the divisor d = 2*div, then n/d = (n * M) >> (31 + l) for all
// n < 2^31 with l = ceil(log2 d), M = ceil(2^(31+l) / d)  (Granlund & Montgomery, N = 31).
//
// Two cheaper modes, chosen at compile time, fold the constants into the accumulator's START value (the first
// dp4a adds it for free): MODE 0, div == 1: start = bias, result = acc.  MODE 1, div = 2^m (m >= 1):
// floor((2*acc + div) / (2*div)) + bias = (acc + div/2 + bias*div) >> m with an arithmetic shift, so
// start = div/2 + bias*div and the result is one shift.  MODE 2 is the general multiply-high form.
struct ConvRound {
    uint32_t M, shift;  // shift = l - 1, applied to the high word of n * M
    int32_t add, K, bias;  // n = 2*acc + add, add = div + d*K
    int32_t mode, start, m;
    template <int MODE>
    __device__ __forceinline__ int32_t quotient(int32_t acc) const  // before the 0..255 clamp
    {
        if (MODE == 0) return acc;
        if (MODE == 1) return acc >> m;
        uint32_t n = (uint32_t)(2 * acc + add);
        return (int32_t)(__umulhi(n, M) >> shift) - K + bias;
    }
    // four results clamped to 0..255 and packed, result 0 in the low byte: two I2IP instructions
    template <int MODE>
    __device__ __forceinline__ uint32_t pack4(int32_t a0, int32_t a1, int32_t a2, int32_t a3) const
    {
        if (MODE == 3)  // coefficients pre-scaled by 256 / div, start 128: the quotient IS byte 1 of the sum (it can not leave 0..255)
            return __byte_perm(__byte_perm((uint32_t)a0, (uint32_t)a1, 0x0051), __byte_perm((uint32_t)a2, (uint32_t)a3, 0x0051), 0x5410);
        if (MODE == 4)  // pre-scaled by 65536 / div, start 32768: the quotient is byte 2
            return __byte_perm(__byte_perm((uint32_t)a0, (uint32_t)a1, 0x0062), __byte_perm((uint32_t)a2, (uint32_t)a3, 0x0062), 0x5410);
        uint32_t hi, out;
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(quotient<MODE>(a3)), "r"(quotient<MODE>(a2)), "r"(0));
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(quotient<MODE>(a1)), "r"(quotient<MODE>(a0)), "r"(hi));
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
    r->mode = 2;
    r->start = 0;
    r->m = 0;
    const int64_t folded = (int64_t)div / 2 + (int64_t)bias * div;  // MODE 1 start value
    if (div == 1 && bias > -(1 << 20) && bias < (1 << 20)) {
        r->mode = 0;
        r->start = bias;
    } else if ((d & (d - 1)) == 0 && folded > -(1ll << 28) && folded < (1ll << 28)) {
        r->mode = 1;
        r->start = (int32_t)folded;
        r->m = l - 1;  // div = 2^(l-1)
    }
    return true;
}

// unsigned pixel bytes times signed coefficient bytes (the CUDA intrinsic has no mixed form)
__device__ __forceinline__ int32_t dp4a_u8s8(uint32_t px4, uint32_t coef4, int32_t acc)
{
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px4), "r"(coef4), "r"(acc));
    return d;
}

// coef = u * v^T with integer factors, v within int8?  (box, binomial/"Gaussian" blurs are; sharpen and
// edge kernels are not)
template <int K>
static bool rank_one(const int32_t *coef, int32_t (&u)[K], int32_t (&v)[K])
{
    int r0 = -1, c0 = -1;
    for (int i = 0; i < K * K && r0 < 0; i++)
        if (coef[i]) { r0 = i / K; c0 = i % K; }
    if (r0 < 0) return false;
    // v = row r0 divided by the gcd of its entries, so that u stays integral whenever a factorisation exists
    int64_t g = 0;
    for (int x = 0; x < K; x++) {
        int64_t a = coef[r0 * K + x] < 0 ? -(int64_t)coef[r0 * K + x] : coef[r0 * K + x], b = g;
        while (b) { int64_t t = a % b; a = b; b = t; }
        g = a;
    }
    for (int x = 0; x < K; x++) {
        v[x] = (int32_t)(coef[r0 * K + x] / g);
        if (v[x] < -128 || v[x] > 127) return false;
    }
    for (int y = 0; y < K; y++) {
        if (coef[y * K + c0] % v[c0]) return false;
        u[y] = coef[y * K + c0] / v[c0];
        for (int x = 0; x < K; x++)
            if ((int64_t)u[y] * v[x] != coef[y * K + x]) return false;
    }
    return true;
}

// 3x3 strip kernels: per tap column dx = -1, 0, +1 the coefficient bytes for the upper / lower row of an output pair
struct Conv3Coef {
    uint32_t a[3], b[3];
    uint32_t ah[3], bh[3];  // WIDE kernels only: coefficients beyond a signed byte are split c = 128 * hi + lo
};
// ppmx_conv_ua.cu: the 3x3 strip kernel at any width / pointer alignment (mode: 0..3 as ConvRound::pack4)
cudaError_t conv3_ua_launch(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const Conv3Coef &cf, const ConvRound &rnd,
                            int mode, bool wide, cudaStream_t s);

// ppmx_conv_sep.cu: 5x5 / 7x7 strip kernels on vertical words (rank-1 with 16-bit column sums, and dense).  Both return
// false when the kernel does not apply (the caller falls through to the older kernels), true with *err set otherwise.
// rh = rows per strip (0 = default).
bool conv_sep16(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div, int32_t bias,
                const ConvRound &rnd, int rh, cudaStream_t s, cudaError_t *err);
bool conv_dense_strip(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, const ConvRound &rnd,
                      int rh, cudaStream_t s, cudaError_t *err);

// ppmx_conv_vw.cu: dense k x k for k = 9, 11, 13, 15 with signed-byte coefficients, vertical words in shared memory
// ... and rank-1 non-negative kernels of those sizes whose column sums fit 16 bits (Gaussian / binomial blurs)
bool conv_vwsep(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div, int32_t bias,
                const ConvRound &rnd, cudaStream_t s, cudaError_t *err);
bool conv_vw(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, const ConvRound &rnd, cudaStream_t s,
             cudaError_t *err);

}  // namespace ppmx
