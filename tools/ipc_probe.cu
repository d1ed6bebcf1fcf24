// Probe: does CUDA IPC peer mapping work between processes on this box, and what is
// the in-kernel flag ping-pong latency / pull bandwidth over NVLink?  (dev tool)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/wait.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("rank %d: %s -> %s\n", rank, #x, cudaGetErrorString(e)); exit(2);} } while (0)

__global__ void pingpong(volatile unsigned long long *mine, volatile unsigned long long *peer, int rank, int rounds)
{
    for (int r = 1; r <= rounds; r++) {
        if (rank == 0) {
            *peer = r; __threadfence_system();
            while (*mine < (unsigned long long)r) {}
        } else {
            while (*mine < (unsigned long long)r) {}
            *peer = r; __threadfence_system();
        }
    }
}
__global__ void pull(const double *peer, double *dst, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[i] = __ldcv(peer + i);
}
int main()
{
    int p01[2], p10[2];
    pipe(p01); pipe(p10);
    pid_t pid = fork();
    int rank = pid == 0 ? 1 : 0;
    CK(cudaSetDevice(rank));
    char *buf; size_t bytes = 1 << 20;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    cudaIpcMemHandle_t h, hp;
    CK(cudaIpcGetMemHandle(&h, buf));
    int wfd = rank == 0 ? p01[1] : p10[1], rfd = rank == 0 ? p10[0] : p01[0];
    write(wfd, &h, sizeof h);
    read(rfd, &hp, sizeof hp);
    char *peer;
    CK(cudaIpcOpenMemHandle((void **)&peer, hp, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaDeviceSynchronize());
    char c = 1; write(wfd, &c, 1); read(rfd, &c, 1); // both mapped
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rounds = 2000;
    cudaEventRecord(e0);
    pingpong<<<1, 1>>>((volatile unsigned long long *)buf, (volatile unsigned long long *)peer, rank, rounds);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rank %d: ping-pong %d rounds %.3f ms -> %.2f us round trip\n", rank, rounds, ms, ms * 1000 / rounds);
    double *dst; CK(cudaMalloc(&dst, 65536 + 8));
    for (int blocks = 1; blocks <= 16; blocks *= 4) {
        pull<<<blocks, 1024>>>((const double *)(peer + 4096), dst, 8193);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int k = 0; k < 50; k++) pull<<<blocks, 1024>>>((const double *)(peer + 4096), dst, 8193);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        printf("rank %d: pull 64 KiB with %d CTAs: %.2f us per kernel (incl. launch)\n", rank, blocks, ms * 1000 / 50);
    }
    c = 1; write(wfd, &c, 1); read(rfd, &c, 1);
    cudaIpcCloseMemHandle(peer);
    if (rank == 0) { int st; wait(&st); printf("child exit %d\n", WEXITSTATUS(st)); }
    return 0;
}
