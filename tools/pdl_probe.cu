// Dev probe: can a 16-CTA cluster kernel (one CTA per SM, ~200 KB of shared memory each, as
// k_wpanel) run beside a machine-filling kernel (2 CTAs per SM, as k_flush_w)?  Three
// orders are timed: serial (plain launches), programmatic dependent launch (the cluster kernel
// executes griddepcontrol.launch_dependents when it starts, the big kernel carries the
// programmatic-stream-serialization attribute and never waits on the grid dependency), and two
// streams with the cluster kernel on a high-priority stream launched after the big one.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(544) k_cluster(unsigned long long spin_ns, unsigned long long *stamp, int trigger)
{
    extern __shared__ double smem[];
    if (trigger) asm volatile("griddepcontrol.launch_dependents;");
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) smem[0] = (double)t0;
    while (gtime() - t0 < spin_ns) {}
    if (threadIdx.x == 0) {
        stamp[2 * blockIdx.x] = t0;
        stamp[2 * blockIdx.x + 1] = gtime();
    }
}

__global__ void __launch_bounds__(64) k_big(unsigned long long spin_ns, unsigned long long *stamp)
{
    extern __shared__ double smem[];
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) smem[0] = (double)t0;
    while (gtime() - t0 < spin_ns) {}
    if (threadIdx.x == 0) {
        stamp[2 * blockIdx.x] = t0;
        stamp[2 * blockIdx.x + 1] = gtime();
    }
}

static void span(const unsigned long long *h, int n, unsigned long long &lo, unsigned long long &hi)
{
    lo = ~0ull, hi = 0;
    for (int i = 0; i < n; i++) {
        if (h[2 * i] < lo) lo = h[2 * i];
        if (h[2 * i + 1] > hi) hi = h[2 * i + 1];
    }
}

int main()
{
    const int NC = 16, NBIG = 148 * 2 * 12; // 12 waves of 2 CTAs per SM
    const unsigned long long cl_ns = 180000, big_ns = 30000;
    const int cl_smem = 200 * 1024, big_smem = 100 * 1024;
    unsigned long long *sc, *sb, *hc, *hb;
    cudaMalloc(&sc, 2 * NC * 8);
    cudaMalloc(&sb, 2 * NBIG * 8);
    hc = (unsigned long long *)malloc(2 * NC * 8);
    hb = (unsigned long long *)malloc(2 * NBIG * 8);
    cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, cl_smem);
    cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
    int lo_p, hi_p;
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    cudaStream_t s, s2;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi_p);
    cudaEvent_t e0, e1, ev;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);

    auto launch_cluster = [&](cudaStream_t st, int trigger) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(NC);
        cfg.blockDim = dim3(544);
        cfg.dynamicSmemBytes = cl_smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = NC;
        at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_cluster, cl_ns, sc, trigger);
    };
    auto launch_big = [&](cudaStream_t st, int pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(NBIG);
        cfg.blockDim = dim3(64);
        cfg.dynamicSmemBytes = big_smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        return cudaLaunchKernelEx(&cfg, k_big, big_ns, sb);
    };
    auto report = [&](const char *name, float ms) {
        cudaMemcpy(hc, sc, 2 * NC * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hb, sb, 2 * NBIG * 8, cudaMemcpyDeviceToHost);
        unsigned long long c0, c1, b0, b1;
        span(hc, NC, c0, c1);
        span(hb, NBIG, b0, b1);
        const unsigned long long z = c0 < b0 ? c0 : b0;
        printf("%-28s total %.3f ms; cluster [%.1f, %.1f] us, big [%.1f, %.1f] us  (%s)\n", name, ms, (c0 - z) * 1e-3,
               (c1 - z) * 1e-3, (b0 - z) * 1e-3, (b1 - z) * 1e-3, cudaGetErrorString(cudaGetLastError()));
    };
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        // serial
        cudaEventRecord(e0, s);
        launch_cluster(s, 0);
        launch_big(s, 0);
        cudaEventRecord(e1, s);
        cudaStreamSynchronize(s);
        cudaEventElapsedTime(&ms, e0, e1);
        report("serial", ms);
        // programmatic dependent launch
        cudaEventRecord(e0, s);
        launch_cluster(s, 1);
        launch_big(s, 1);
        cudaEventRecord(e1, s);
        cudaStreamSynchronize(s);
        cudaEventElapsedTime(&ms, e0, e1);
        report("pdl (cluster first)", ms);
        // two streams, big first, cluster on the high-priority stream
        cudaEventRecord(e0, s);
        cudaEventRecord(ev, s);
        cudaStreamWaitEvent(s2, ev, 0);
        launch_big(s, 0);
        launch_cluster(s2, 0);
        cudaEventRecord(ev, s2);
        cudaStreamWaitEvent(s, ev, 0);
        cudaEventRecord(e1, s);
        cudaStreamSynchronize(s);
        cudaEventElapsedTime(&ms, e0, e1);
        report("2 streams, big first", ms);
        // two streams, cluster first
        cudaEventRecord(e0, s);
        cudaEventRecord(ev, s);
        cudaStreamWaitEvent(s2, ev, 0);
        launch_cluster(s2, 0);
        launch_big(s, 0);
        cudaEventRecord(ev, s2);
        cudaStreamWaitEvent(s, ev, 0);
        cudaEventRecord(e1, s);
        cudaStreamSynchronize(s);
        cudaEventElapsedTime(&ms, e0, e1);
        report("2 streams, cluster first", ms);
    }
    return 0;
}
