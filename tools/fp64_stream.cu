// Dev tool: the rank-t update as a pure streaming kernel -- k_flush_t's main loop (operands
// out of shared memory, fixed P / F tiles) over a real 8192 x 16384 tableau with the tile loads
// prefetched one tile ahead and the results stored, nothing else (no multiplier staging, no
// pivot-row marks, no block bookkeeping).  Shows what the inner loop + HBM streaming can reach.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_stream fp64_stream.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROWS = 64, LANES = 128, HALVES = 2, TC = 2 * LANES;

template <int TR, int PF, int SYNC, int FG> // rows per thread tile, prefetch depth, __syncthreads every SYNC tiles (0: never), FG: multipliers straight from global (L1) instead of shared
__global__ void __launch_bounds__(256, 2) k(double *tab, int m, int Cl, int t, const double *gF)
{
    extern __shared__ double sm[];
    double *sP = sm, *sF = sm + (size_t)32 * TC;
    const int tid = threadIdx.x, lane = tid % LANES, half = tid / LANES;
    for (int e = tid; e < 32 * TC; e += 256) sP[e] = 1e-9 * (e % 97);
    for (int e = tid; e < 32 * ROWS; e += 256) sF[e] = 1e-7 * (e % 13);
    __syncthreads();
    const double *sPl = sP + 2 * lane;
    constexpr int TPB = ROWS / (HALVES * TR);
    const int nrb = m / ROWS, ctiles = Cl / TC;
    const long long units = (long long)nrb * ctiles;
    const long long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    const long long tiles = (u1 - u0) * TPB;
    auto addr = [&](long long tile, int w) {
        const long long u = u0 + tile / TPB;
        const int k = (int)(tile % TPB);
        const int ct = (int)(u / nrb), rb = (int)(u % nrb) * ROWS;
        return tab + (size_t)(rb + (k * HALVES + half) * TR + w) * Cl + ct * TC + 2 * lane;
    };
    double2 nx[PF][TR];
#pragma unroll
    for (int pf = 0; pf < PF; pf++)
        if (pf < tiles)
#pragma unroll
            for (int w = 0; w < TR; w++) nx[pf][w] = *reinterpret_cast<const double2 *>(addr(pf, w));
    for (long long tile = 0; tile < tiles; tile++) {
        double2 a[TR];
#pragma unroll
        for (int w = 0; w < TR; w++) a[w] = nx[0][w];
#pragma unroll
        for (int pf = 0; pf + 1 < PF; pf++)
#pragma unroll
            for (int w = 0; w < TR; w++) nx[pf][w] = nx[pf + 1][w];
        if (tile + PF < tiles)
#pragma unroll
            for (int w = 0; w < TR; w++) nx[PF - 1][w] = *reinterpret_cast<const double2 *>(addr(tile + PF, w));
        const int rc = ((int)(tile % TPB) * HALVES + half) * TR;
        const int rbase = (int)((u0 + tile / TPB) % nrb) * ROWS;
#pragma unroll 8
        for (int s = 0; s < t; s++) {
            const double2 p2 = *reinterpret_cast<const double2 *>(sPl + (size_t)s * TC);
            const double *f = FG ? gF + (size_t)s * m + rbase + rc : sF + (size_t)s * ROWS + rc;
#pragma unroll
            for (int w = 0; w < TR; w += 2) {
                const double2 f2 = FG ? __ldg(reinterpret_cast<const double2 *>(f + w)) : *reinterpret_cast<const double2 *>(f + w);
                a[w].x = __dadd_rn(a[w].x, __dmul_rn(f2.x, p2.x));
                a[w].y = __dadd_rn(a[w].y, __dmul_rn(f2.x, p2.y));
                a[w + 1].x = __dadd_rn(a[w + 1].x, __dmul_rn(f2.y, p2.x));
                a[w + 1].y = __dadd_rn(a[w + 1].y, __dmul_rn(f2.y, p2.y));
            }
        }
#pragma unroll
        for (int w = 0; w < TR; w++) *reinterpret_cast<double2 *>(addr(tile, w)) = a[w];
        if (SYNC > 0 && (tile + 1) % SYNC == 0) __syncthreads(); // what staging shared operands per unit costs
    }
}

template <int TR, int PF, int SYNC, int FG>
void run(const char *name, int sm, double *tab, int m, int Cl, int t, const double *gF)
{
    const size_t smem = ((size_t)32 * TC + 32 * ROWS) * 8;
    cudaFuncSetAttribute(k<TR, PF, SYNC, FG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<TR, PF, SYNC, FG><<<sm * 2, 256, smem>>>(tab, m, Cl, t, gF);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaError_t e = cudaGetLastError();
    const double inst = 2.0 * m * Cl * t;
    printf("%-34s t=%2d: %7.1f us  FP64 %.0f%% of 18.54 T/s  HBM %.0f GB/s  %s\n", name, t, ms * 1e3,
           100.0 * inst / ms / 1e9 / 18.54, 2.0 * m * Cl * 8 / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const int m = 8192, Cl = 16384;
    double *tab;
    cudaMalloc(&tab, (size_t)m * Cl * 8);
    cudaMemset(tab, 0, (size_t)m * Cl * 8);
    double *gF;
    cudaMalloc(&gF, (size_t)32 * m * 8);
    cudaMemset(gF, 0, (size_t)32 * m * 8);
    for (int t : {32, 24, 8}) {
        run<8, 1, 0, 0>("8x2, pf 1, F in smem, no barrier", sm, tab, m, Cl, t, gF);
        run<8, 1, 4, 0>("8x2, pf 1, F in smem, barrier/unit", sm, tab, m, Cl, t, gF);
        run<8, 1, 0, 1>("8x2, pf 1, F from global (L1)", sm, tab, m, Cl, t, gF);
        run<8, 2, 0, 1>("8x2, pf 2, F from global (L1)", sm, tab, m, Cl, t, gF);
    }
    return 0;
}
