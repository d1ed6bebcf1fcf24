// Dev tool: measured FP64 pipe throughput on this GPU, fused (DFMA) and non-fused
// (DMUL + DADD, what the reference's rounding order forces on the update kernels).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double f, double p)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) a[i] = __fma_rn(f, p, a[i]);
            else a[i] = __dadd_rn(a[i], __dmul_rn(f, a[(i + 1) & 15] * 0 + p)); // keep the mul live per element
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) k2(double *out, int iters, const double *fp)
{ // non-fused with distinct operands per element, like the flush: a = a + f_w * p_c
    double a[16], f[8], p[2];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = fp[i];
    p[0] = fp[8];
    p[1] = fp[9];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int w = 0; w < 8; w++) {
            a[2 * w] = __dadd_rn(a[2 * w], __dmul_rn(f[w], p[0]));
            a[2 * w + 1] = __dadd_rn(a[2 * w + 1], __dmul_rn(f[w], p[1]));
        }
        p[0] += 1e-30; // defeat hoisting of the products
        p[1] -= 1e-30;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sm * 8, iters = 20000;
    double *out, *fp, h[10] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996, 1e-9, -1e-9};
    cudaMalloc(&out, (size_t)grid * 256 * 8);
    cudaMalloc(&fp, 80);
    cudaMemcpy(fp, h, 80, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
            else k2<1><<<grid, 256>>>(out, iters, fp);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double inst = (double)grid * 256 * iters * 16 * (mode == 0 ? 1 : 2) + (mode ? (double)grid * 256 * iters * 2 : 0);
        printf("%s: %.3f ms, %.2f T FP64 instructions(lanes)/s, per SM per clk at 1.965 GHz: %.1f lanes\n",
               mode == 0 ? "DFMA (fused)" : "DMUL+DADD (non-fused)", ms, inst / ms / 1e9, inst / (ms * 1e-3) / sm / 1.965e9);
    }
    return 0;
}
