// Context management and small shared host helpers of the C ABI.
#include "xp_common.cuh"

extern "C" const char *xp_version(void) { return "xpoly_b200 0.1 (sm_100a)"; }

extern "C" int xp_ctx_create(int device, xp_ctx **out)
{
    if (!out) return XP_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    xp_ctx *ctx = new xp_ctx();
    ctx->device = device;
    *out = ctx; // returned even on failure so xp_last_error() can explain
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        ctx->err = std::string("no usable CUDA device: ") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                   " (xpoly_b200 has no CPU fallback)";
        return XP_ERR_CUDA;
    }
    XP_CUDA_OK(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    XP_CUDA_OK(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    XP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    XP_CUDA_OK(ctx, cudaEventCreate(&ctx->ev0));
    XP_CUDA_OK(ctx, cudaEventCreate(&ctx->ev1));
    return 0;
}

extern "C" void xp_ctx_destroy(xp_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->stream) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        xp_large_release_cached(ctx);
        if (ctx->scratch) cudaFree(ctx->scratch);
        if (ctx->stage) cudaFreeHost(ctx->stage);
        if (ctx->gws) cudaFree(ctx->gws);
        if (ctx->pipe_copy) {
            cudaStreamDestroy(ctx->pipe_copy);
            cudaEventDestroy(ctx->pipe_begin);
            for (int c = 0; c < XP_PIPE_MAX; c++) {
                cudaStreamDestroy(ctx->pipe_stream[c]);
                cudaEventDestroy(ctx->pipe_up[c]);
                cudaEventDestroy(ctx->pipe_done[c]);
            }
        }
        cudaEventDestroy(ctx->ev0);
        cudaEventDestroy(ctx->ev1);
        cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

extern "C" const char *xp_last_error(const xp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" uint64_t xp_ctx_launch_count(const xp_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" float xp_ctx_last_kernel_ms(const xp_ctx *ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }
extern "C" void *xp_ctx_stream(const xp_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int xp_ctx_scratch(xp_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) {
            XP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            XP_CUDA_OK(ctx, cudaFree(ctx->scratch));
            ctx->scratch = nullptr;
            ctx->scratch_bytes = 0;
        }
        size_t want = bytes + (bytes >> 2) + 256;
        XP_CUDA_OK(ctx, cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return 0;
}

int xp_ctx_gws(xp_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->gws_bytes) {
        if (ctx->gws) {
            XP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            XP_CUDA_OK(ctx, cudaFree(ctx->gws));
            ctx->gws = nullptr;
            ctx->gws_bytes = 0;
        }
        XP_CUDA_OK(ctx, cudaMalloc(&ctx->gws, bytes));
        ctx->gws_bytes = bytes;
    }
    *out = ctx->gws;
    return 0;
}

int xp_ctx_pipe(xp_ctx *ctx)
{
    if (ctx->pipe_copy) return 0;
    XP_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->pipe_begin, cudaEventDisableTiming));
    for (int c = 0; c < XP_PIPE_MAX; c++) {
        XP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->pipe_stream[c], cudaStreamNonBlocking));
        XP_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->pipe_up[c], cudaEventDisableTiming));
        XP_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->pipe_done[c], cudaEventDisableTiming));
    }
    XP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->pipe_copy, cudaStreamNonBlocking)); // last: marks "ready"
    return 0;
}

// Pinned host buffers for callers that want full-rate H2D/D2H on the large path.
extern "C" int xp_host_alloc(xp_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    XP_CUDA_OK(ctx, cudaMallocHost(out, bytes));
    return 0;
}

extern "C" int xp_host_free(xp_ctx *ctx, void *p)
{
    if (!ctx) return XP_ERR_BAD_ARG;
    XP_CUDA_OK(ctx, cudaFreeHost(p));
    return 0;
}
