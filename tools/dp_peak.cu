// tools/dp_peak.cu -- measures the non-fused FP64 issue rate of this GPU: independent chains of
// DMUL + DADD (never contracted to DFMA), the instruction mix of the bicubic operators.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/dp_peak tools/dp_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) dp_kernel(double *out, double a, double b, int iters)
{
    double v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = a + i + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) v[i] = __dadd_rn(__dmul_rn(v[i], b), a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += v[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b, int iters)
{
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = a + i + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fma(v[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 12345.678) out[0] = s;
}

// I2F.F64.U8 with byte selectors (the u8 -> double conversion of the bicubic kernels), alone and
// interleaved 1:2 with DMUL/DADD as in the resize inner loop
__global__ void __launch_bounds__(256) i2f_kernel(double *out, const uchar4 *in, int iters)
{
    uchar4 v = in[threadIdx.x & 31];
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int it = 0; it < iters; it++) {
        s0 += (double)v.x;
        s1 += (double)v.y;
        s2 += (double)v.z;
        s3 += (double)v.w;
        v.x += 1; v.y += 3; v.z += 5; v.w += 7;
    }
    if (s0 + s1 + s2 + s3 == 12345.678) out[0] = s0;
}

__global__ void __launch_bounds__(256) mix_kernel(double *out, const uchar4 *in, double w, int iters)
{
    uchar4 v = in[threadIdx.x & 31];
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int it = 0; it < iters; it++) {
        s0 = __dadd_rn(s0, __dmul_rn((double)v.x, w));
        s1 = __dadd_rn(s1, __dmul_rn((double)v.y, w));
        s2 = __dadd_rn(s2, __dmul_rn((double)v.z, w));
        s3 = __dadd_rn(s3, __dmul_rn((double)v.w, w));
        v.x += 1; v.y += 3; v.z += 5; v.w += 7;
    }
    if (s0 + s1 + s2 + s3 == 12345.678) out[0] = s0;
}

__global__ void __launch_bounds__(256) mix_trick_kernel(double *out, const uchar4 *in, double w, int iters)
{
    uchar4 v = in[threadIdx.x & 31];
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int it = 0; it < iters; it++) {
        s0 = __dadd_rn(s0, __dmul_rn(__hiloint2double(0x43300000, v.x) - 4503599627370496.0, w));
        s1 = __dadd_rn(s1, __dmul_rn(__hiloint2double(0x43300000, v.y) - 4503599627370496.0, w));
        s2 = __dadd_rn(s2, __dmul_rn(__hiloint2double(0x43300000, v.z) - 4503599627370496.0, w));
        s3 = __dadd_rn(s3, __dmul_rn(__hiloint2double(0x43300000, v.w) - 4503599627370496.0, w));
        v.x += 1; v.y += 3; v.z += 5; v.w += 7;
    }
    if (s0 + s1 + s2 + s3 == 12345.678) out[0] = s0;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double *d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * 8;
    for (int rep = 0; rep < 3; rep++) {
        dp_kernel<8><<<grid, 256>>>(d, 1.0000001, 0.9999999, iters);
        cudaEventRecord(e0);
        dp_kernel<8><<<grid, 256>>>(d, 1.0000001, 0.9999999, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)grid * 256 * iters * 8 * 2;  // DMUL + DADD
        printf("{\"kind\": \"dmul+dadd\", \"dp_inst_per_s\": %.4g, \"per_sm_per_clk_at_max\": %.2f, \"ms\": %.3f}\n", ops / (ms * 1e-3),
               ops / (ms * 1e-3) / sms / (clk_khz * 1e3), ms);
        dfma_kernel<<<grid, 256>>>(d, 1.0000001, 0.9999999, iters);
        cudaEventRecord(e0);
        dfma_kernel<<<grid, 256>>>(d, 1.0000001, 0.9999999, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        ops = (double)grid * 256 * iters * 8;
        printf("{\"kind\": \"dfma\", \"dp_inst_per_s\": %.4g, \"per_sm_per_clk_at_max\": %.2f, \"ms\": %.3f}\n", ops / (ms * 1e-3),
               ops / (ms * 1e-3) / sms / (clk_khz * 1e3), ms);
    }
    {
        uchar4 *din;
        cudaMalloc(&din, 32 * sizeof(uchar4));
        cudaMemset(din, 7, 32 * sizeof(uchar4));
        float ms;
        for (int which = 0; which < 3; which++) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (which == 0) i2f_kernel<<<grid, 256>>>(d, din, iters);
                else if (which == 1) mix_kernel<<<grid, 256>>>(d, din, 0.9999, iters);
                else mix_trick_kernel<<<grid, 256>>>(d, din, 0.9999, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            double conv = (double)grid * 256 * iters * 4;
            printf("{\"kind\": \"%s\", \"conversions_per_s\": %.4g, \"per_sm_per_clk_at_max\": %.2f, \"ms\": %.3f}\n",
                   which == 0 ? "i2f.f64.u8 + dadd" : which == 1 ? "i2f + dmul + dadd (resize inner loop)" : "2^52 trick + dmul + dadd",
                   conv / (ms * 1e-3), conv / (ms * 1e-3) / sms / (clk_khz * 1e3), ms);
        }
    }
    printf("{\"sms\": %d, \"max_clock_mhz\": %d}\n", sms, clk_khz / 1000);
    return 0;
}
