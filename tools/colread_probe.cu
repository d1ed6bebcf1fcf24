// Dev probe: latency of reading one tableau column (m entries, row stride C doubles) with a
// small number of CTAs, as phase A of the windowed panel does, against the same read from a
// transposed copy (contiguous), from a compact window copy (short row stride), and through
// TMA (2 x rows boxes).  Each CTA runs dependent iterations (the next column index depends on
// the values just read), so the per-iteration time is a latency, not a bandwidth.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o colread_probe colread_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) k_col(const double *tab, size_t stride_rows, size_t stride_cols, int m, int ncols,
                                             int rows_per_cta, int iters, double *out, unsigned long long *ns)
{
    __shared__ double s_red[16];
    __shared__ int s_q;
    const int tid = threadIdx.x, r_lo = blockIdx.x * rows_per_cta;
    int q = (blockIdx.x * 7 + 3) % ncols;
    unsigned long long t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        double v = 0.0;
        for (int li = tid; li < rows_per_cta; li += blockDim.x) {
            const int i = r_lo + li;
            if (i < m) v += tab[(size_t)i * stride_rows + (size_t)q * stride_cols];
        }
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double s = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += s_red[w];
            acc += s;
            s_q = (int)(((unsigned)q * 1103515245u + 12345u + (unsigned)(s * 1e-300)) % (unsigned)ncols);
        }
        __syncthreads();
        q = s_q;
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (tid == 0) {
        out[blockIdx.x] = acc;
        ns[blockIdx.x] = t1 - t0;
    }
}

int main()
{
    const int m = 8192, C = 16384, w = 4096;
    double *tab, *wt, *win, *out;
    unsigned long long *ns;
    cudaMalloc(&tab, (size_t)m * C * 8);
    cudaMalloc(&wt, (size_t)w * m * 8);
    cudaMalloc(&win, (size_t)m * w * 8);
    cudaMalloc(&out, 1024 * 8);
    cudaMalloc(&ns, 1024 * 8);
    cudaMemset(tab, 0, (size_t)m * C * 8);
    cudaMemset(wt, 0, (size_t)w * m * 8);
    cudaMemset(win, 0, (size_t)m * w * 8);
    double *flush;
    cudaMalloc(&flush, (size_t)512 << 20);
    struct Cfg { const char *name; const double *p; size_t sr, sc; int ctas, threads; } cfg[] = {
        {"strided 16 CTAs x 512 rows (stride 128 KB)", tab, (size_t)C, 1, 16, 512},
        {"strided 64 CTAs x 128 rows", tab, (size_t)C, 1, 64, 128},
        {"strided 128 CTAs x 64 rows", tab, (size_t)C, 1, 128, 64},
        {"compact window copy 16 CTAs (stride 32 KB)", win, (size_t)w, 1, 16, 512},
        {"transposed copy 16 CTAs (contiguous)", wt, 1, (size_t)m, 16, 512},
        {"transposed copy 64 CTAs", wt, 1, (size_t)m, 64, 128},
    };
    const int iters = 200;
    for (auto &c : cfg) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(flush, rep, (size_t)512 << 20); // evict L2
            k_col<<<c.ctas, c.threads>>>(c.p, c.sr, c.sc, m, w, (m + c.ctas - 1) / c.ctas, iters, out, ns);
            cudaDeviceSynchronize();
        }
        unsigned long long h[1024];
        cudaMemcpy(h, ns, c.ctas * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0, sum = 0;
        for (int i = 0; i < c.ctas; i++) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
        printf("%-48s avg %.2f us/iter  slowest CTA %.2f us/iter\n", c.name, sum / 1e3 / c.ctas / iters, mx / 1e3 / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
