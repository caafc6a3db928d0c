// ppmx_conv_sep.cu -- EXTENSION (no reference counterpart, parity unpinned): 5x5 / 7x7 integer convolutions on "vertical words".
// Part of libppmx_gpu.so; conventions in ppmx_common.cuh, rounding in ppmx_conv.cuh.
//
// Layout shared by both kernels (it is the box kernel's): a thread owns 16 byte columns (one 16-byte vector per row) and walks
// down a strip of RH rows TWO rows at a time; a warp covers 30 chunks of a row, lanes 0 and 31 only supply their neighbours
// (so every writing lane has both neighbours in its own warp and the halo columns come by shuffle, no shared memory, no
// barrier).  Per byte column the thread keeps the 8 source rows around the output pair (y, y+1) as two words
//     W_lo = rows y-3 .. y      W_hi = rows y+1 .. y+4          (byte b of a word = one row)
// Moving down two rows costs 1.25 PRMT per byte (one 2x4 transpose of the two new rows, then two byte-window moves per column).
// A 7-tap (or 5-tap) vertical dot product of a column is then two dp4a against the column factor laid out in the bytes of the
// rows it multiplies; the lower output row of the pair uses the same two words with the factor moved up one byte.
//
// rank-1 kernels, coef = u v^T (binomial / "Gaussian" blurs, Sobel-like derivatives): `conv_sep16_kernel`.
//   The column sums S fit 16 bits (checked on the host: 255 * sum|u|), so two of them pack into one register,
//   Q[c] = S[c] | S[c+3] << 16 (byte columns 3 apart = horizontal neighbours of one channel), and the horizontal pass is
//   (K-1)/2 dp2a (two taps each) + one multiply-add: per output byte at K = 7
//       fma pipe: 2 IDP.4A + 3 IDP.2A + 1 IMAD        alu pipe: 1.25 PRMT (window) + 1 PRMT (pairing) + 0.75 PRMT (quotient)
//   against 2 IDP + 4 IMAD / 2 PRMT + 3 IADD3 + 1.5 finishing for the 32-bit form this replaces (0.47 of the HBM roofline).
//   Normalised non-negative blurs with a power-of-two divisor have u and v pre-scaled so that the total scale is 65536 / div:
//   the rounded quotient is then byte 2 of the sum (start value 32768) -- no shift, no saturation.
// dense kernels (rank > 1, e.g. emboss / unsharp / LoG masks): `conv_dense_strip_kernel`.
//   Output (y, c) = sum over the K tap columns j of the two dp4a of column c + 3 (j - R) against that tap column's coefficients:
//   2K IDP per byte (14 at 7x7, where the old planar kernel needed 17.5 plus a de-interleaving pass through shared memory);
//   the neighbours' words come by shuffle.  This is the fma pipe's bound for a dense k x k on 8-bit dot products (aligned
//   4-row words leave 1 of 8 multipliers idle at 7 taps): 14 / 64 lanes per clock per SM.
#include "ppmx_conv.cuh"

namespace ppmx {


template <bool SIGNED>
__device__ __forceinline__ int32_t vdot4(uint32_t px4, uint32_t coef4, int32_t acc)
{
    int32_t d;
    if (SIGNED) asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px4), "r"(coef4), "r"(acc));
    else asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(px4), "r"(coef4), "r"(acc));
    return d;
}
// two 16-bit column sums times two coefficient bytes (bytes 0,1 of b: lo; bytes 2,3: hi)
template <bool SIGNED, bool HI>
__device__ __forceinline__ int32_t hdot2(uint32_t s2, uint32_t coef, int32_t acc)
{
    int32_t d;
    if (SIGNED) {
        if (HI) asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(s2), "r"(coef), "r"(acc));
        else asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(s2), "r"(coef), "r"(acc));
    } else {
        if (HI) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(s2), "r"(coef), "r"(acc));
        else asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(s2), "r"(coef), "r"(acc));
    }
    return d;
}

// ---- the sliding 8-row window of a thread's 16 byte columns -----------------------------------------------------------
struct VWindow {
    uint32_t lo[16], hi[16];
    // rows i0..i5 = the six rows above the first pair's two new rows: hi = rows i2..i5, lo bytes 2,3 = rows i0, i1
    __device__ __forceinline__ void init(const uint4 (&f)[6])
    {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t r0 = (&f[0].x)[w], r1 = (&f[1].x)[w], r2 = (&f[2].x)[w], r3 = (&f[3].x)[w], r4 = (&f[4].x)[w],
                           r5 = (&f[5].x)[w];
            const uint32_t a = __byte_perm(r0, r1, 0x5140), b = __byte_perm(r0, r1, 0x7362);  // [r0.0 r1.0 r0.1 r1.1], [.2 .2 .3 .3]
            const uint32_t c = __byte_perm(r2, r3, 0x5140), d = __byte_perm(r2, r3, 0x7362);
            const uint32_t e = __byte_perm(r4, r5, 0x5140), g = __byte_perm(r4, r5, 0x7362);
            lo[4 * w + 0] = a << 16;
            lo[4 * w + 1] = a & 0xffff0000u;
            lo[4 * w + 2] = b << 16;
            lo[4 * w + 3] = b & 0xffff0000u;
            hi[4 * w + 0] = __byte_perm(c, e, 0x5410);
            hi[4 * w + 1] = __byte_perm(c, e, 0x7632);
            hi[4 * w + 2] = __byte_perm(d, g, 0x5410);
            hi[4 * w + 3] = __byte_perm(d, g, 0x7632);
        }
    }
    // two rows further down: rows A, B enter at the top bytes of hi
    __device__ __forceinline__ void slide(const uint4 &A, const uint4 &B)
    {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t ra = (&A.x)[w], rb = (&B.x)[w];
            const uint32_t tl = __byte_perm(ra, rb, 0x5140), th = __byte_perm(ra, rb, 0x7362);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = 4 * w + j;
                lo[c] = __byte_perm(lo[c], hi[c], 0x5432);
                hi[c] = __byte_perm(hi[c], j < 2 ? tl : th, (j & 1) ? 0x7632 : 0x5432);
            }
        }
    }
};

// column factor u (K taps) laid out in the bytes of the 8-row window: for the upper output row of the pair tap i multiplies
// window row i - R + 3, for the lower one window row i - R + 4
template <int K>
static void window_coef(const int32_t (&u)[K], uint32_t &a_lo, uint32_t &a_hi, uint32_t &b_lo, uint32_t &b_hi)
{
    constexpr int R = K / 2;
    a_lo = a_hi = b_lo = b_hi = 0;
    for (int i = 0; i < K; i++) {
        const uint32_t byte = (uint32_t)(uint8_t)u[i];
        const int pa = i - R + 3, pb = i - R + 4;
        if (pa < 4) a_lo |= byte << (8 * pa);
        else a_hi |= byte << (8 * (pa - 4));
        if (pb < 4) b_lo |= byte << (8 * pb);
        else b_hi |= byte << (8 * (pb - 4));
    }
}

// =========================================================================================================================
// rank-1 kernel with 16-bit column sums
// =========================================================================================================================
template <int K>
struct Sep16Coef {
    uint32_t ua_lo, ua_hi, ub_lo, ub_hi;
    uint32_t v[2];   // v[0 .. K-2] as bytes, tap 2p, 2p+1 = byte pair p
    int32_t v_last;  // v[K-1]
    int32_t centre;  // CENTRE kernels: coef = u v^T + centre * delta (unsharp masks: (1 + k) div delta - k G)
};

template <int K, int MODE, bool SIGNED, bool CENTRE, int RH, bool INNER, bool EDGE, int PF = 2>
__device__ __forceinline__ void conv_sep16_body(const RowSource &rs, uint8_t *__restrict__ dst, uint32_t nchunks, int chunk, int ys,
                                                const Sep16Coef<K> &cf, const ConvRound &rnd)
{
    constexpr int R = K / 2, H = 3 * R, SH = H, QR = H > 6 ? H - 6 : 0, NP = (K - 1) / 2;
    static_assert(RH % (2 * PF) == 0, "strip height");
    const int lane = threadIdx.x & 31;
    const bool valid = chunk >= 0 && chunk < (int)nchunks;
    const uint32_t cx = (uint32_t)(chunk < 0 ? 0 : chunk >= (int)nchunks ? (int)nchunks - 1 : chunk);
    const bool writes = valid && lane >= 1 && lane <= 30;
    const bool left = EDGE && chunk == 0, right = EDGE && chunk == (int)nchunks - 1;
    const size_t pitch = (size_t)nchunks * 16;
    const int gy0 = rs.y0 + ys;
    const uint8_t *src = rs.own + (size_t)cx * 16 + (size_t)(INNER ? ys - 3 : 0) * pitch;
    auto load_row = [&](int i) {  // row i counted from the window's first row (gy0 - 3)
        const uint8_t *p = INNER ? src + (size_t)i * pitch : rs.row(gy0 - 3 + i, pitch) + (size_t)cx * 16;
        return __ldg(reinterpret_cast<const uint4 *>(p));
    };
    pdl_wait();
    VWindow W;
    uint4 nb[PF][2];
    {
        uint4 f[6];
#pragma unroll
        for (int i = 0; i < 6; i++) f[i] = load_row(i);
#pragma unroll
        for (int u = 0; u < PF; u++) {
            nb[u][0] = load_row(6 + 2 * u);
            nb[u][1] = load_row(7 + 2 * u);
        }
        W.init(f);
    }

    // cw / csel: the window words and the byte of them that hold the output row itself (the centre tap of CENTRE kernels)
    auto emit = [&](uint32_t ulo, uint32_t uhi, uint8_t *o, const uint32_t(&cw)[16], uint32_t csel) {
        int32_t S[16 + SH];  // column sums of byte columns 0 .. 15 + SH
#pragma unroll
        for (int c = 0; c < 16; c++) S[c] = vdot4<SIGNED>(W.hi[c], uhi, vdot4<SIGNED>(W.lo[c], ulo, 0));
#pragma unroll
        for (int i = 0; i < SH; i++) S[16 + i] = __shfl_down_sync(0xffffffffu, S[i], 1);
        // pixel W-1+m mirrors to pixel W-m: column 16+i comes from column 13 - 3 (i/3) + i%3
        auto Smr = [&](int i) { return (uint32_t)S[13 - 3 * (i / 3) + i % 3]; };
        if (EDGE) {
#pragma unroll
            for (int i = 0; i < SH; i++) S[16 + i] = right ? (int32_t)Smr(i) : S[16 + i];
        }
        // the left neighbour's last H column sums too, in the same shuffle round (packing the halo pairs here costs H + QR more
        // PRMT than shuffling the neighbours' packed pairs would, but a second, dependent round of shuffles is gone)
        int32_t SL[H];
#pragma unroll
        for (int i = 0; i < H; i++) SL[i] = __shfl_up_sync(0xffffffffu, S[16 - H + i], 1);
        if (EDGE) {  // pixel -m mirrors to pixel m-1: column -3m + ch comes from column 3 (m-1) + ch
#pragma unroll
            for (int i = 0; i < H; i++) SL[i] = left ? S[3 * (R - 1 - i / 3) + i % 3] : SL[i];
        }
        auto Sx = [&](int c) { return (uint32_t)(c >= 0 ? S[c] : SL[c + H]); };
        uint32_t Q[H + 16 + QR];  // Q[H + c] = S[c] | S[c+3] << 16 for c = -H .. 15 + QR
#pragma unroll
        for (int c = -H; c < 16 + QR; c++) Q[H + c] = __byte_perm(Sx(c), Sx(c + 3), 0x5410);
        uint32_t ov[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int32_t a[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = 4 * b + j;
                int32_t t = CENTRE ? rnd.start + cf.centre * (int32_t)__byte_perm(cw[c], 0u, csel) : rnd.start;
#pragma unroll
                for (int p = 0; p < NP; p++) {  // taps 2p, 2p+1 at byte columns c + 3 (2p - R), c + 3 (2p + 1 - R)
                    const uint32_t q = Q[H + c + 3 * (2 * p - R)];
                    t = (p & 1) ? hdot2<SIGNED, true>(q, cf.v[p >> 1], t) : hdot2<SIGNED, false>(q, cf.v[p >> 1], t);
                }
                a[j] = t + cf.v_last * S[c + H];
            }
            ov[b] = rnd.template pack4<MODE>(a[0], a[1], a[2], a[3]);
        }
        if (writes) *reinterpret_cast<uint4 *>(o) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    };

    uint8_t *out = dst + (size_t)ys * pitch + (size_t)cx * 16;
#pragma unroll 1
    for (int g0 = 0; g0 < RH / 2; g0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int g = g0 + u;
            if (!INNER && ys + 2 * g >= rs.h) return;
            const uint4 A = nb[u][0], B = nb[u][1];
            if (g + PF < RH / 2 && (INNER || ys + 2 * (g + PF) < rs.h)) {
                nb[u][0] = load_row(6 + 2 * (g + PF));
                nb[u][1] = load_row(7 + 2 * (g + PF));
            }
            W.slide(A, B);
            emit(cf.ua_lo, cf.ua_hi, out, W.lo, 0x4443u);  // row y = byte 3 of W_lo
            if (INNER || ys + 2 * g + 1 < rs.h) emit(cf.ub_lo, cf.ub_hi, out + pitch, W.hi, 0x4440u);  // row y+1 = byte 0 of W_hi
            out += 2 * pitch;
        }
    }
}

template <int K, int MODE, bool SIGNED, bool CENTRE, int RH, int MINB>
__global__ void __launch_bounds__(128, MINB) conv_sep16_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t nchunks,
                                                            const Sep16Coef<K> cf, const ConvRound rnd)
{
    pdl_trigger();
    const int wg = blockIdx.x * 4 + (threadIdx.x >> 5);  // warp number along the row: 30 chunks each
    if (wg * 30 >= (int)nchunks) return;
    const int chunk = wg * 30 - 1 + (int)(threadIdx.x & 31);
    const int ys = blockIdx.y * RH;
    const bool inner = ys >= 3 && ys + RH + 3 <= rs.h;  // window rows ys-3 .. ys+RH+2 all in the own band
    const bool edge = wg == 0 || wg * 30 + 30 >= (int)nchunks;
    if (inner && !edge) conv_sep16_body<K, MODE, SIGNED, CENTRE, RH, true, false, (MINB >= 4 ? 1 : 2)>(rs, dst, nchunks, chunk, ys, cf, rnd);
    else if (inner) conv_sep16_body<K, MODE, SIGNED, CENTRE, RH, true, true, (MINB >= 4 ? 1 : 2)>(rs, dst, nchunks, chunk, ys, cf, rnd);
    else conv_sep16_body<K, MODE, SIGNED, CENTRE, RH, false, true, (MINB >= 4 ? 1 : 2)>(rs, dst, nchunks, chunk, ys, cf, rnd);
}

template <int K, int RH, int MINB>
static cudaError_t conv_sep16_launch(const RowSource &rs, uint8_t *dst, uint32_t nchunks, uint32_t h, const Sep16Coef<K> &cf,
                                     const ConvRound &rnd, int mode, bool sgn, cudaStream_t s)
{
    dim3 grid((nchunks + 119) / 120, (h + RH - 1) / RH);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
#define PPMX_SEP16(MODE, SGN, CEN) launch(conv_sep16_kernel<K, MODE, SGN, CEN, RH, MINB>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, rnd)
    if (cf.centre) {  // (always the signed form)
        if (mode == 0) PPMX_SEP16(0, true, true);
        else if (mode == 1) PPMX_SEP16(1, true, true);
        else PPMX_SEP16(2, true, true);
    } else if (sgn) {
        if (mode == 0) PPMX_SEP16(0, true, false);
        else if (mode == 1) PPMX_SEP16(1, true, false);
        else PPMX_SEP16(2, true, false);
    } else {
        if (mode == 4) PPMX_SEP16(4, false, false);
        else if (mode == 0) PPMX_SEP16(0, false, false);
        else if (mode == 1) PPMX_SEP16(1, false, false);
        else PPMX_SEP16(2, false, false);
    }
#undef PPMX_SEP16
    return PPMX_LAUNCHED();
}

// coef = u v^T + delta at the centre element only (unsharp masks, "identity plus blur")?  The rank-1 part is fixed by any row and
// column that avoid the centre.
template <int K>
static bool rank_one_centre(const int32_t *coef, int32_t (&u)[K], int32_t (&v)[K], int32_t *delta)
{
    constexpr int R = K / 2;
    for (int r0 = 0; r0 < K; r0++) {
        if (r0 == R) continue;
        int c0 = -1;
        for (int x = 0; x < K && c0 < 0; x++)
            if (x != R && coef[r0 * K + x]) c0 = x;
        if (c0 < 0) continue;
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
        }
        for (int y = 0; y < K; y++)
            for (int x = 0; x < K; x++)
                if (!(y == R && x == R) && (int64_t)u[y] * v[x] != coef[y * K + x]) return false;
        const int64_t d = (int64_t)coef[R * K + R] - (int64_t)u[R] * v[R];
        if (d < -(1ll << 24) || d > (1ll << 24)) return false;
        *delta = (int32_t)d;
        return true;
    }
    return false;
}

template <int K>
static bool conv_sep16_k(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef, int32_t div, int32_t bias,
                         ConvRound rnd, int rh, cudaStream_t s, cudaError_t *err)
{
    int32_t u[K], v[K], centre = 0;
    if (!rank_one<K>(coef, u, v) && !rank_one_centre<K>(coef, u, v, &centre)) return false;
    bool upos = true, uneg = true, vpos = true, vneg = true;
    int64_t su = 0, sabs = 0, sv = 0;
    for (int i = 0; i < K; i++) {
        upos = upos && u[i] >= 0, uneg = uneg && u[i] <= 0, vpos = vpos && v[i] >= 0, vneg = vneg && v[i] <= 0;
        sabs += u[i] < 0 ? -(int64_t)u[i] : u[i];
    }
    if (uneg && vneg) {  // (-u)(-v)^T is the same kernel
        for (int i = 0; i < K; i++) u[i] = -u[i], v[i] = -v[i];
        upos = vpos = true;
    }
    for (int i = 0; i < K; i++) su += u[i], sv += v[i];
    int mode = rnd.mode;
    bool sgn = true;
    if (upos && vpos && 255 * su <= 65535 && !centre) {
        sgn = false;
        for (int i = 0; i < K; i++)
            if (u[i] > 255 || v[i] > 255) return false;
        // normalised blur with a power-of-two divisor: scale u and v so that the quotient is byte 2 of the sum
        if (mode == 1 && bias == 0 && div >= 1 && div <= 65536 && (div & (div - 1)) == 0 && su * sv == div) {
            const int64_t scale = 65536 / div;
            for (int64_t a = scale; a >= 1; a >>= 1) {  // the largest share of the scale the column factor can take
                const int64_t b = scale / a;
                bool ok = 255 * su * a <= 65535;
                for (int i = 0; i < K; i++) ok = ok && u[i] * a <= 255 && v[i] * b <= 255;
                if (ok) {
                    for (int i = 0; i < K; i++) u[i] = (int32_t)(u[i] * a), v[i] = (int32_t)(v[i] * b);
                    rnd.start = 32768;
                    mode = 4;
                    break;
                }
            }
        }
    } else {
        if (255 * sabs > 32767) return false;
        for (int i = 0; i < K; i++)
            if (u[i] < -128 || u[i] > 127 || v[i] < -128 || v[i] > 127) return false;
    }
    Sep16Coef<K> cf;
    window_coef<K>(u, cf.ua_lo, cf.ua_hi, cf.ub_lo, cf.ub_hi);
    cf.v[0] = cf.v[1] = 0;
    for (int i = 0; i < K - 1; i++) cf.v[i >> 2] |= (uint32_t)(uint8_t)v[i] << (8 * (i & 3));
    cf.v_last = v[K - 1];
    cf.centre = centre;
    const uint32_t nchunks = w * 3 / 16;
    // rows per strip / CTAs per SM (register budget) / row pairs prefetched.  7x7 binomial at 8192^2, fraction of the HBM roofline:
    // 16 / 4 / 1 (116 registers, no spills: the default) 0.600; 16 / 3 / 2 (168 registers) 0.570; 32 / 4 / 1 0.600; 16 / 5 / 1
    // (96 registers) 0.571; with two pairs prefetched 4 CTAs per SM spill (0.561).  5x5: 0.666 / 0.624 / 0.654 / 0.664.  The
    // schedulers had 2.8 warps each and one eligible in 1.28 of them per cycle at 3 CTAs per SM: occupancy, not the mix.
    // Measured and dropped: the odd tap as a shift-add, and the three middle taps of a symmetric v as one dp2a of
    // (S[c-3] + S[c+3], S[c]) -- ptxas issues the adds as IMAD.IADD on the same fma pipe, no change (0.561 vs 0.567).
#ifdef PPMX_TUNING
    if (rh == 1) *err = conv_sep16_launch<K, 16, 3>(rs, dst, nchunks, h, cf, rnd, mode, sgn, s);
    else if (rh == 2) *err = conv_sep16_launch<K, 32, 4>(rs, dst, nchunks, h, cf, rnd, mode, sgn, s);
    else if (rh == 3) *err = conv_sep16_launch<K, 32, 3>(rs, dst, nchunks, h, cf, rnd, mode, sgn, s);
    else if (rh == 4) *err = conv_sep16_launch<K, 64, 4>(rs, dst, nchunks, h, cf, rnd, mode, sgn, s);
    else if (rh == 5) *err = conv_sep16_launch<K, 16, 5>(rs, dst, nchunks, h, cf, rnd, mode, sgn, s);
    else
#endif
    *err = conv_sep16_launch<K, 16, 4>(rs, dst, nchunks, h, cf, rnd, mode, sgn, s);
    return true;
}

bool conv_sep16(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, int32_t div, int32_t bias,
                const ConvRound &rnd, int rh, cudaStream_t s, cudaError_t *err)
{
    if (k == 5) return conv_sep16_k<5>(rs, dst, w, h, coef, div, bias, rnd, rh, s, err);
    if (k == 7) return conv_sep16_k<7>(rs, dst, w, h, coef, div, bias, rnd, rh, s, err);
    return false;
}

// =========================================================================================================================
// dense kernel: K tap columns, two dp4a each
// =========================================================================================================================
template <int K>
struct DenseCoef {
    uint32_t a_lo[K], a_hi[K], b_lo[K], b_hi[K];  // per tap column: the coefficient column in the window's bytes (upper / lower row)
};

// WIDE: coefficients beyond a signed byte (-16320 .. 16319) are split c = 128 * hi + lo: a second chain of dot products, weighted 128
template <int K, int MODE, bool WIDE, int RH, bool INNER, bool EDGE, int PF>
__device__ __forceinline__ void conv_dense_body(const RowSource &rs, uint8_t *__restrict__ dst, uint32_t nchunks, int chunk, int ys,
                                                const DenseCoef<K> &cf, const DenseCoef<K> &cfh, const ConvRound &rnd)
{
    constexpr int R = K / 2, H = 3 * R;
    static_assert(RH % (2 * PF) == 0, "strip height");
    const int lane = threadIdx.x & 31;
    const bool valid = chunk >= 0 && chunk < (int)nchunks;
    const uint32_t cx = (uint32_t)(chunk < 0 ? 0 : chunk >= (int)nchunks ? (int)nchunks - 1 : chunk);
    const bool writes = valid && lane >= 1 && lane <= 30;
    const bool left = EDGE && chunk == 0, right = EDGE && chunk == (int)nchunks - 1;
    const size_t pitch = (size_t)nchunks * 16;
    const int gy0 = rs.y0 + ys;
    const uint8_t *src = rs.own + (size_t)cx * 16 + (size_t)(INNER ? ys - 3 : 0) * pitch;
    auto load_row = [&](int i) {
        const uint8_t *p = INNER ? src + (size_t)i * pitch : rs.row(gy0 - 3 + i, pitch) + (size_t)cx * 16;
        return __ldg(reinterpret_cast<const uint4 *>(p));
    };
    pdl_wait();
    VWindow W;
    uint4 nb[PF][2];
    {
        uint4 f[6];
#pragma unroll
        for (int i = 0; i < 6; i++) f[i] = load_row(i);
#pragma unroll
        for (int u = 0; u < PF; u++) {
            nb[u][0] = load_row(6 + 2 * u);
            nb[u][1] = load_row(7 + 2 * u);
        }
        W.init(f);
    }
    uint8_t *out = dst + (size_t)ys * pitch + (size_t)cx * 16;
#pragma unroll 1
    for (int g0 = 0; g0 < RH / 2; g0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int g = g0 + u;
            if (!INNER && ys + 2 * g >= rs.h) return;
            const uint4 A = nb[u][0], B = nb[u][1];
            if (g + PF < RH / 2 && (INNER || ys + 2 * (g + PF) < rs.h)) {
                nb[u][0] = load_row(6 + 2 * (g + PF));
                nb[u][1] = load_row(7 + 2 * (g + PF));
            }
            W.slide(A, B);
            // the neighbours' words of byte columns -H .. -1 and 16 .. 15 + H
            uint32_t Llo[H], Lhi[H], Rlo[H], Rhi[H];
#pragma unroll
            for (int i = 0; i < H; i++) {
                Llo[i] = __shfl_up_sync(0xffffffffu, W.lo[16 - H + i], 1);
                Lhi[i] = __shfl_up_sync(0xffffffffu, W.hi[16 - H + i], 1);
                Rlo[i] = __shfl_down_sync(0xffffffffu, W.lo[i], 1);
                Rhi[i] = __shfl_down_sync(0xffffffffu, W.hi[i], 1);
            }
            if (EDGE) {  // mirror at the raster's left / right edge (same column map as the rank-1 kernel)
#pragma unroll
                for (int i = 0; i < H; i++) {
                    const int ml = 3 * (R - 1 - i / 3) + i % 3, mr = 13 - 3 * (i / 3) + i % 3;
                    Llo[i] = left ? W.lo[ml] : Llo[i];
                    Lhi[i] = left ? W.hi[ml] : Lhi[i];
                    Rlo[i] = right ? W.lo[mr] : Rlo[i];
                    Rhi[i] = right ? W.hi[mr] : Rhi[i];
                }
            }
            auto XL = [&](int c) { return c < 0 ? Llo[c + H] : c < 16 ? W.lo[c] : Rlo[c - 16]; };
            auto XH = [&](int c) { return c < 0 ? Lhi[c + H] : c < 16 ? W.hi[c] : Rhi[c - 16]; };
            const bool second = INNER || ys + 2 * g + 1 < rs.h;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t ov[4];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    int32_t a[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int c = 4 * b + j;
                        int32_t t = rnd.start, th = 0;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const int cc = c + 3 * (k - R);
                            t = vdot4<true>(XL(cc), half ? cf.b_lo[k] : cf.a_lo[k], t);
                            t = vdot4<true>(XH(cc), half ? cf.b_hi[k] : cf.a_hi[k], t);
                            if (WIDE) {
                                th = vdot4<true>(XL(cc), half ? cfh.b_lo[k] : cfh.a_lo[k], th);
                                th = vdot4<true>(XH(cc), half ? cfh.b_hi[k] : cfh.a_hi[k], th);
                            }
                        }
                        a[j] = WIDE ? t + (th << 7) : t;
                    }
                    ov[b] = rnd.template pack4<MODE>(a[0], a[1], a[2], a[3]);
                }
                if (writes && (half == 0 || second)) *reinterpret_cast<uint4 *>(out + (half ? pitch : 0)) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
            }
            out += 2 * pitch;
        }
    }
}

template <int K, int MODE, bool WIDE, int RH, int MINB>
__global__ void __launch_bounds__(128, MINB) conv_dense_strip_kernel(RowSource rs, uint8_t *__restrict__ dst, uint32_t nchunks,
                                                                  const DenseCoef<K> cf, const DenseCoef<K> cfh, const ConvRound rnd)
{
    pdl_trigger();
    const int wg = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (wg * 30 >= (int)nchunks) return;
    const int chunk = wg * 30 - 1 + (int)(threadIdx.x & 31);
    const int ys = blockIdx.y * RH;
    const bool inner = ys >= 3 && ys + RH + 3 <= rs.h;
    const bool edge = wg == 0 || wg * 30 + 30 >= (int)nchunks;
    if (inner && !edge) conv_dense_body<K, MODE, WIDE, RH, true, false, (MINB >= 4 ? 1 : 2)>(rs, dst, nchunks, chunk, ys, cf, cfh, rnd);
    else if (inner) conv_dense_body<K, MODE, WIDE, RH, true, true, (MINB >= 4 ? 1 : 2)>(rs, dst, nchunks, chunk, ys, cf, cfh, rnd);
    else conv_dense_body<K, MODE, WIDE, RH, false, true, (MINB >= 4 ? 1 : 2)>(rs, dst, nchunks, chunk, ys, cf, cfh, rnd);
}

template <int K, int RH, int MINB>
static cudaError_t conv_dense_launch(const RowSource &rs, uint8_t *dst, uint32_t nchunks, uint32_t h, const DenseCoef<K> &cf,
                                     const DenseCoef<K> &cfh, bool wide, const ConvRound &rnd, cudaStream_t s)
{
    dim3 grid((nchunks + 119) / 120, (h + RH - 1) / RH);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
#define PPMX_DENSE(MODE)                                                                                                   \
    do {                                                                                                                   \
        if (wide) launch(conv_dense_strip_kernel<K, MODE, true, RH, MINB>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, cfh, rnd);    \
        else launch(conv_dense_strip_kernel<K, MODE, false, RH, MINB>, grid, dim3(128), 0, s, rs, dst, nchunks, cf, cfh, rnd);        \
    } while (0)
    if (rnd.mode == 0) PPMX_DENSE(0);
    else if (rnd.mode == 1) PPMX_DENSE(1);
    else PPMX_DENSE(2);
#undef PPMX_DENSE
    return PPMX_LAUNCHED();
}

template <int K>
static bool conv_dense_k(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, const int32_t *coef, const ConvRound &rnd, int rh,
                         cudaStream_t s, cudaError_t *err)
{
    DenseCoef<K> cf, cfh;
    bool wide = false;
    for (int k = 0; k < K; k++) {
        int32_t lo[K], hi[K];
        for (int i = 0; i < K; i++) {  // c = 128 * hi + lo with lo in -64..63; hi == 0 for every byte-sized coefficient
            const int32_t v = coef[i * K + k];
            if (v < -16320 || v > 16319) return false;  // (16320 would need hi = 128)
            lo[i] = v, hi[i] = 0;
            if (v < -128 || v > 127) {
                lo[i] = ((v + 64) & 127) - 64;
                hi[i] = (v - lo[i]) / 128;
                wide = true;
            }
        }
        window_coef<K>(lo, cf.a_lo[k], cf.a_hi[k], cf.b_lo[k], cf.b_hi[k]);
        window_coef<K>(hi, cfh.a_lo[k], cfh.a_hi[k], cfh.b_lo[k], cfh.b_hi[k]);
    }
    const uint32_t nchunks = w * 3 / 16;
    // 4 CTAs per SM with one row pair prefetched (121 / 110 registers at 7x7 / 5x5) against 3 with two (variant: rh code 1)
#ifdef PPMX_TUNING
    if (rh == 1) *err = conv_dense_launch<K, 16, 3>(rs, dst, nchunks, h, cf, cfh, wide, rnd, s);
    else if (rh == 2 || rh == 3) *err = conv_dense_launch<K, 32, 4>(rs, dst, nchunks, h, cf, cfh, wide, rnd, s);
    else if (rh == 4) *err = conv_dense_launch<K, 64, 4>(rs, dst, nchunks, h, cf, cfh, wide, rnd, s);
    else
#endif
    *err = conv_dense_launch<K, 16, 4>(rs, dst, nchunks, h, cf, cfh, wide, rnd, s);
    return true;
}

bool conv_dense_strip(const RowSource &rs, uint8_t *dst, uint32_t w, uint32_t h, int k, const int32_t *coef, const ConvRound &rnd,
                      int rh, cudaStream_t s, cudaError_t *err)
{
    if (k == 5) return conv_dense_k<5>(rs, dst, w, h, coef, rnd, rh, s, err);
    if (k == 7) return conv_dense_k<7>(rs, dst, w, h, coef, rnd, rh, s, err);
    return false;
}

}  // namespace ppmx
