// Dev tool: how fast can the rank-t update's inner loop run when its operands come out of
// shared memory the way k_flush_t reads them (P: one lane-distinct 128-bit load per two
// columns, F: broadcast 128-bit loads, two rows each), for several register tile shapes.
// No global traffic: this isolates FP64 pipe vs shared-memory operand delivery.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_tile fp64_tile.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int TR, int TCOL> // rows x columns per thread; 128 lanes x 2 halves
__global__ void __launch_bounds__(256, 2) k(double *out, int tiles, int t)
{
    extern __shared__ double sm[];
    constexpr int LANES = 128, ROWS = 64;
    const int TC = TCOL * LANES;
    double *sP = sm, *sF = sm + (size_t)32 * TC;
    const int tid = threadIdx.x, lane = tid % LANES, half = tid / LANES;
    for (int e = tid; e < 32 * TC; e += 256) sP[e] = 1e-9 * (e % 97);
    for (int e = tid; e < 32 * ROWS; e += 256) sF[e] = 1.0 + 1e-7 * (e % 13);
    __syncthreads();
    const double *sPl = sP + TCOL * lane;
    double acc = 0;
    for (int tile = 0; tile < tiles; tile++) {
        double a[TR][TCOL];
#pragma unroll
        for (int w = 0; w < TR; w++)
#pragma unroll
            for (int c = 0; c < TCOL; c++) a[w][c] = tile + w + c;
        const int rc = ((tile * 2 + half) * TR) % (ROWS - TR + 1) & ~1;
#pragma unroll 8
        for (int s = 0; s < t; s++) {
            double p[TCOL];
#pragma unroll
            for (int c = 0; c < TCOL; c += 2) {
                const double2 p2 = *reinterpret_cast<const double2 *>(sPl + (size_t)s * TC + c);
                p[c] = p2.x;
                p[c + 1] = p2.y;
            }
            const double *f = sF + (size_t)s * ROWS + rc;
#pragma unroll
            for (int w = 0; w < TR; w += 2) {
                const double2 f2 = *reinterpret_cast<const double2 *>(f + w);
#pragma unroll
                for (int c = 0; c < TCOL; c++) {
                    a[w][c] = __dadd_rn(a[w][c], __dmul_rn(f2.x, p[c]));
                    a[w + 1][c] = __dadd_rn(a[w + 1][c], __dmul_rn(f2.y, p[c]));
                }
            }
        }
#pragma unroll
        for (int w = 0; w < TR; w++)
#pragma unroll
            for (int c = 0; c < TCOL; c++) acc += a[w][c];
    }
    out[blockIdx.x * 256 + tid] = acc;
}

template <int TR, int TCOL>
void run(const char *name, int sm, double *out)
{
    const int grid = sm * 2, tiles = 400, t = 32;
    const size_t smem = ((size_t)32 * TCOL * 128 + 32 * 64) * 8;
    cudaFuncSetAttribute(k<TR, TCOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k<TR, TCOL><<<grid, 256, smem>>>(out, tiles, t);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaError_t e = cudaGetLastError();
    const double inst = (double)grid * 256 * tiles * t * TR * TCOL * 2;
    printf("%s: %.3f ms  %.2f T FP64 lanes/s (%.0f%% of 18.54)  smem %zu B  %s\n", name, ms, inst / ms / 1e9,
           100.0 * inst / ms / 1e9 / 18.54, smem, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, (size_t)sm * 2 * 256 * 8);
    run<8, 2>("8 rows x 2 cols (k_flush_t)", sm, out);
    run<4, 4>("4 rows x 4 cols", sm, out);
    run<8, 4>("8 rows x 4 cols", sm, out);
    run<16, 2>("16 rows x 2 cols", sm, out);
    run<4, 2>("4 rows x 2 cols", sm, out);
    return 0;
}
