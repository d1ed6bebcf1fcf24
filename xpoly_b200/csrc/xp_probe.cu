// Measurement probe: the non-fused FP64 throughput of this GPU, in-run.
//
// The update kernels of this library are bound by the FP64 pipe issuing separate DMUL and DADD
// (the reference rounds the product and the sum separately, lpsol.h:1487-1488, so DFMA is not
// allowed).  bench.py's roofline divides by what this probe reaches on the device it runs on:
// 16 independent accumulators per thread, a = a + f * p with operands in registers, nothing else
// in the loop.
#include "xp_common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_probe_fp64(double *out, int iters, const double *fp)
{
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
            a[2 * w] = xp_add(a[2 * w], xp_mul(f[w], p[0]));
            a[2 * w + 1] = xp_add(a[2 * w + 1], xp_mul(f[w], p[1]));
        }
        p[0] = xp_add(p[0], 1e-30); // keeps the products inside the loop
        p[1] = xp_add(p[1], -1e-30);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace

// *tops: 1e12 FP64 operations per second (one DMUL or one DADD of one lane = one operation).
extern "C" int xp_probe_fp64_nonfused(xp_ctx *ctx, double *tops, double *ms_out)
{
    if (!ctx || !tops) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int grid = ctx->sm_count * 8, iters = 12000;
    void *scr = nullptr;
    int rc = xp_ctx_scratch(ctx, (size_t)grid * 256 * 8 + 128, &scr);
    if (rc) return rc;
    double *out = (double *)scr, *fp = out + (size_t)grid * 256;
    const double h[10] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003,
                          0.9999997, 1.0000004, 0.9999996, 1e-9,      -1e-9};
    cudaStream_t s = ctx->stream;
    XP_CUDA_OK(ctx, cudaMemcpyAsync(fp, h, sizeof h, cudaMemcpyHostToDevice, s));
    float best = 0.f;
    for (int rep = 0; rep < 3; rep++) { // first launch warms up; keep the best of the rest
        XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, s));
        k_probe_fp64<<<grid, 256, 0, s>>>(out, iters, fp);
        XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, s));
        XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
        float ms = 0.f;
        XP_CUDA_OK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
        ctx->launches++;
    }
    const double ops = (double)grid * 256 * iters * (16 * 2 + 2);
    *tops = ops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}
