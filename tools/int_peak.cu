// tools/int_peak.cu -- measures the issue rate of the integer instructions the convolution and colour kernels are
// made of (IDP.4A, IMAD, PRMT, SHF, IADD3, LOP3, I2IP) alone and in 1:1 mixes, to learn which of them share a pipe.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/int_peak tools/int_peak.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum { OP_IDP, OP_IMAD, OP_PRMT, OP_SHF, OP_IADD3, OP_LOP3, OP_I2IP, OP_IDP2A, OP_NONE };

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t v, uint32_t a, uint32_t b)
{
    uint32_t d = v;
    if (OP == OP_IDP) asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(v));
    if (OP == OP_IDP2A) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(v));
    if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(a), "r"(b));
    if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(a), "r"(b));
    if (OP == OP_SHF) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(a), "r"(b));
    if (OP == OP_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(d) : "r"(v), "r"(a), "r"(b));
    if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(v), "r"(a), "r"(b));
    if (OP == OP_I2IP) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(a), "r"(b));
    return d;
}

// CH independent chains per op; each iteration issues CH x OPA and CH x OPB
template <int OPA, int OPB, int CH>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t a, uint32_t b, int iters)
{
    uint32_t va[CH], vb[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        va[i] = a * (i + 1) + threadIdx.x;
        vb[i] = b * (i + 3) + threadIdx.x;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) {
                va[i] = step<OPA>(va[i], a, b);
                if (OPB != OP_NONE) vb[i] = step<OPB>(vb[i], a, b);
            }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += va[i] ^ vb[i];
    if (s == 0x12345679u) out[0] = s;
}

static int sms, clk_khz;
static uint32_t *d;

template <int OPA, int OPB>
static void run(const char *name)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2048, grid = sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<OPA, OPB, 6><<<grid, 256>>>(d, 0x01020304u + rep, 0x3210u, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double n = (double)grid * 256 * iters * 4 * 6 * (OPB == OP_NONE ? 1 : 2);
    printf("{\"mix\": \"%s\", \"thread_inst_per_s\": %.4g, \"per_sm_per_clk_at_max\": %.1f, \"ms\": %.3f}\n", name,
           n / (best * 1e-3), n / (best * 1e-3) / sms / (clk_khz * 1e3), best);
}

int main()
{
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    cudaMalloc(&d, 8);
    run<OP_IDP, OP_NONE>("idp4a");
    run<OP_IDP2A, OP_NONE>("idp2a");
    run<OP_IMAD, OP_NONE>("imad");
    run<OP_PRMT, OP_NONE>("prmt");
    run<OP_SHF, OP_NONE>("shf");
    run<OP_IADD3, OP_NONE>("iadd3");
    run<OP_LOP3, OP_NONE>("lop3");
    run<OP_I2IP, OP_NONE>("i2ip");
    run<OP_IDP, OP_IMAD>("idp4a+imad");
    run<OP_IDP, OP_PRMT>("idp4a+prmt");
    run<OP_IDP, OP_IADD3>("idp4a+iadd3");
    run<OP_IDP, OP_SHF>("idp4a+shf");
    run<OP_IDP, OP_I2IP>("idp4a+i2ip");
    run<OP_IMAD, OP_IADD3>("imad+iadd3");
    run<OP_IMAD, OP_PRMT>("imad+prmt");
    run<OP_PRMT, OP_IADD3>("prmt+iadd3");
    run<OP_PRMT, OP_SHF>("prmt+shf");
    run<OP_I2IP, OP_PRMT>("i2ip+prmt");
    run<OP_I2IP, OP_IMAD>("i2ip+imad");
    return 0;
}
