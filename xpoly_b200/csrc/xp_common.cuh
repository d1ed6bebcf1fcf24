// Shared device/host helpers for the xpoly_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/xpoly_b200.h"

#define XP_EPS 0.00000000000000001 /* INFINITESIMAL, reference flty.h:46 */

// Internal (device-side) status values beyond the public 0..4.
#define XPI_RUNNING (-1)
#define XPI_OPT_PENDING (-2) /* no c_j > 0 left: feasibility check outstanding */

#define XP_PIPE_MAX 16 /* chunks in flight of a pipelined batched host call */

struct xp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    size_t smem_optin = 0;
    uint64_t launches = 0;
    float last_kernel_ms = 0.f;
    std::string err;
    // reusable device scratch
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *cached_lp = nullptr; // xp_lp_f64 reused by xp_six_slack_f64 / xp_six_two_stage_f64_large
    void *stage = nullptr;      // small pinned staging buffer (gathered constant column of an uploaded LP)
    size_t stage_bytes = 0;
    void *cached_aux = nullptr; // the auxiliary (phase 1) handle of xp_six_two_stage_f64_large
    int slack_block = 0;       // pivots per flush for xp_six_slack_f64 (0 = automatic)
    int slack_window = 0;      // pricing window for xp_six_slack_f64 / xp_six_two_stage_f64_large (0 = automatic)
    // global-memory state slabs of the batched kernels (LPs beyond shared memory)
    void *gws = nullptr;
    size_t gws_bytes = 0;
    // upload/solve pipeline of the batched host-pointer entry points (xp_ctx_pipe)
    cudaStream_t pipe_copy = nullptr;
    cudaStream_t pipe_stream[XP_PIPE_MAX] = {};
    cudaEvent_t pipe_begin = nullptr, pipe_up[XP_PIPE_MAX] = {}, pipe_done[XP_PIPE_MAX] = {};
};

#define XP_CUDA_OK(ctx, expr)                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            char b__[512];                                                                     \
            snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
                     cudaGetErrorString(e__));                                                 \
            (ctx)->err = b__;                                                                  \
            return XP_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------
// Float comparison semantics of the reference (flty.cpp:41-95), bit-faithful.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ bool xp_feq(double a, double b)
{
    if ((a > 0 && b < 0) || (a < 0 && b > 0)) return false;
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    if ((a == 0.0 && b <= XP_EPS) || (b == 0.0 && a <= XP_EPS)) return true;
    if (a > b) return (a - b) <= XP_EPS;
    return (b - a) <= XP_EPS;
}
__host__ __device__ __forceinline__ bool xp_fle(double a, double b) { return a < b || xp_feq(a, b); }
__host__ __device__ __forceinline__ bool xp_fge(double a, double b) { return a > b || xp_feq(a, b); }

#ifdef __CUDACC__
// IEEE mul / add with one rounding each and no FMA contraction, as the
// reference's `a + f*x` compiles without -march (lpsol.h:1487-1488).
__device__ __forceinline__ double xp_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xp_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xp_div(double a, double b) { return __ddiv_rn(a, b); }

// Matrix::mulOfRow / Matrix::mul scalar short-circuits (matt.h:1335-1341,
// :1358-1364): v == 1 leaves x untouched, v == 0 (tolerant) zeroes it.
__device__ __forceinline__ double xp_scale(double x, double v, bool v_is_one, bool v_is_zero)
{
    return v_is_one ? x : (v_is_zero ? 0.0 : xp_mul(x, v));
}

// ---- (value, index) arg-min with the reference's tie rule: the first strict
// minimum in index order == lowest index among IEEE-equal minima. ----
struct XpMinIdx {
    double v;
    int i; // -1 = empty
};
__device__ __forceinline__ XpMinIdx xp_better(XpMinIdx a, XpMinIdx b)
{
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v < a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ XpMinIdx xp_warp_argmin(XpMinIdx x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        XpMinIdx y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = xp_better(x, y);
    }
    return x;
}
__device__ __forceinline__ int xp_warp_min_int(int x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// Block-wide reductions; `sh` needs 33 slots of the given type.  All threads
// must call; result valid in every thread.
__device__ __forceinline__ XpMinIdx xp_block_argmin(XpMinIdx x, XpMinIdx *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    x = xp_warp_argmin(x);
    __syncthreads();
    if (lane == 0) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        XpMinIdx y;
        y.v = 0.0;
        y.i = -1;
        if (lane < nw) y = sh[lane];
        y = xp_warp_argmin(y);
        if (lane == 0) sh[32] = y;
    }
    __syncthreads();
    return sh[32];
}
__device__ __forceinline__ int xp_block_min_int(int x, int *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    x = xp_warp_min_int(x);
    __syncthreads();
    if (lane == 0) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        int y = lane < nw ? sh[lane] : 0x7fffffff;
        y = xp_warp_min_int(y);
        if (lane == 0) sh[32] = y;
    }
    __syncthreads();
    return sh[32];
}
#endif // __CUDACC__

// host-side helpers shared by the translation units
int xp_ctx_scratch(xp_ctx *ctx, size_t bytes, void **out);
int xp_ctx_gws(xp_ctx *ctx, size_t bytes, void **out);
int xp_ctx_pipe(xp_ctx *ctx); // create the pipeline streams / events on first use
void xp_large_release_cached(xp_ctx *ctx);
