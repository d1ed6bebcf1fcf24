// Batched FP64 simplex, one CTA per LP, tableau in shared memory:
// SIX<FloatMat,Float>::TwoStageMethod for many small independent LPs
// (dependence-feasibility shape, B&B node relaxations).  FP64 arithmetic policy
// for the skeleton in xp_batch_core.cuh; bit-faithful to the reference's Float
// semantics (flty.cpp:41-131) and operation order (lpsol.h:1455-1511).
#include "xp_batch_core.cuh"
#include "xp_batch_warp_f64.cuh"

#include <cstdlib>

#include <cstring>

namespace {

struct OpsF64 {
    typedef double E;
    typedef double In;
    typedef XpMinIdx Key;

    __device__ static __forceinline__ E zero() { return 0.0; }
    __device__ static __forceinline__ E from_int(int i) { return (double)i; }
    __device__ static __forceinline__ E from_in(In x) { return x; }
    __device__ static __forceinline__ bool in_pos(In x) { return x > 0.0; }
    __device__ static __forceinline__ bool in_neg(In x) { return x < 0.0; }
    __device__ static __forceinline__ bool pos(E x) { return x > 0.0; }
    __device__ static __forceinline__ bool le_zero(E x) { return xp_fle(x, 0.0); }
    __device__ static __forceinline__ bool is_zero(E x) { return xp_feq(x, 0.0); }

    __device__ static __forceinline__ Key empty_key()
    {
        Key k;
        k.v = 0.0;
        k.i = -1;
        return k;
    }
    __device__ static __forceinline__ Key make_key(E b, E a, int i)
    {
        Key k;
        k.v = xp_div(b, a); // v = rhs / coeff, lpsol.h:603
        k.i = i;
        return k;
    }
    __device__ static __forceinline__ Key better(Key a, Key b) { return xp_better(a, b); }
    __device__ static __forceinline__ Key shfl_xor(Key x, int o)
    {
        Key y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        return y;
    }
    __device__ static __forceinline__ int key_index(Key k) { return k.i; }

    __device__ static __forceinline__ void bind(XpB<E> &, long long *) {}
    __device__ static __forceinline__ void reset(XpB<E> &) {}

    // Row of the first minimum constant term (lpsol.h:894-904).
    __device__ static int argmin_rhs(XpB<E> &S)
    {
        Key best = empty_key();
        for (int i = threadIdx.x; i < S.m; i += blockDim.x) {
            Key k;
            k.v = S.tab[i * S.LD + S.n];
            k.i = i;
            best = xp_better(best, k);
        }
        best = xpb_block_best<OpsF64>(S, best);
        return best.i;
    }

    // SIX::pivot (lpsol.h:1455-1511) on row p, entering variable q.
    __device__ static int pivot(XpB<E> &S, int p, int q)
    {
        const int tid = threadIdx.x, LD = S.LD, C = S.C, n = S.n, m = S.m;
        const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
        const int bv = S.eq2bv[p];
        const double pv = S.tab[p * LD + q];
        const double cq = S.tgtf[q];
        const double r = xp_div(1.0, pv);
        const bool r_one = xp_feq(r, 1.0), r_zero = xp_feq(r, 0.0);
        const bool cq_zero = xp_feq(cq, 0.0), cq_one = xp_feq(cq, 1.0);
        __syncthreads();
        for (int i = tid; i < m; i += blockDim.x)
            if (i != p) S.fcol[i] = -S.tab[i * LD + q]; // coeff_of_nv = -eq(i, nv), :1485
        double *rowp = S.tab + p * LD;
        for (int j = tid; j < C; j += blockDim.x) rowp[j] = xp_scale(rowp[j], r, r_one, r_zero); // :1471
        __syncthreads();
        for (int j = tid; j < C; j += blockDim.x) { // objective row, :1496-1501
            double t = xp_mul(rowp[j], -1.0);
            if (j >= n) t = -t;
            t = cq_zero ? 0.0 : (cq_one ? t : xp_mul(t, cq));
            S.tgtf[j] = xp_add(t, S.tgtf[j]);
        }
        for (int i = w; i < m; i += nw) { // rank-1 elimination, :1481-1490
            if (i == p) continue;
            const double f = S.fcol[i];
            double *row = S.tab + i * LD;
            for (int j = lane; j < C; j += 32) row[j] = xp_add(row[j], xp_mul(f, rowp[j]));
        }
        if (tid == 0) xpb_swap_basis(S, p, q, bv);
        S.pivots++;
        __syncthreads();
        return 0;
    }

    // Optimal exit: sol from the basis + is_feasible (lpsol.h:1089-1127, :783-822)
    // with vc = -I | 0.
    __device__ static int optimal_exit(XpB<E> &S)
    {
        const int tid = threadIdx.x, LD = S.LD, n = S.n;
        for (int j = tid; j < S.C; j += blockDim.x)
            S.sol[j] = (j < n && !S.nvset[j]) ? S.tab[S.bv2eq[j] * LD + n] : 0.0;
        __syncthreads();
        int bad = 0;
        for (int j = tid; j < n; j += blockDim.x)
            if (xp_mul(-1.0, S.sol[j]) > 0.0) bad = 1; // vc(i,i)*sol(i) > vc(i,rhs)
        for (int i = tid; i < S.m; i += blockDim.x) {
            const double *row = S.tab + i * LD;
            double sum = 0.0; // left-to-right; non-basic terms are exact +-0
            for (int j = 0; j < n; j++)
                if (!S.nvset[j]) sum = xp_add(sum, xp_mul(row[j], S.sol[j]));
            if (!xp_feq(sum, row[n])) bad = 1;
        }
        bad = __syncthreads_or(bad);
        return bad ? XP_SIX_OPTIMAL_IS_INFEASIBLE : XP_SIX_SUCC;
    }

    // lpsol.h:944-953 with FloatMat::substit (xmat.cpp:1491-1520), is_eq=false.
    __device__ static int restore_objective(XpB<E> &S, const In *tg, int n_orig)
    {
        const int tid = threadIdx.x, LD = S.LD, C = S.C, rhs = S.n;
        for (int j = tid; j < C; j += blockDim.x)
            S.tgtf[j] = j < n_orig ? tg[j] : (j == rhs ? tg[n_orig] : 0.0);
        __syncthreads();
        for (int i = 0; i < rhs; i++) {
            const double ci = S.tgtf[i];
            if (xp_feq(ci, 0.0) || S.nvset[i]) continue; // uniform across the CTA
            const double *ex = S.tab + S.bv2eq[i] * LD;
            const double ev = ex[i];
            const bool skip = xp_feq(ev, 0.0);
            double s = -1.0;
            if (!xp_feq(ci, ev)) s = xp_div(-ci, ev);
            const bool s_zero = xp_feq(s, 0.0), s_one = xp_feq(s, 1.0);
            __syncthreads();
            for (int j = tid; j < C; j += blockDim.x) {
                double tj = S.tgtf[j];
                if (j >= rhs) tj = xp_mul(tj, -1.0);
                if (!skip) {
                    double x = s_zero ? 0.0 : (s_one ? ex[j] : xp_mul(ex[j], s));
                    tj = xp_add(x, tj);
                }
                if (j >= rhs) tj = xp_mul(tj, -1.0);
                S.tgtf[j] = tj;
            }
            __syncthreads();
        }
        return 0;
    }

    __device__ static void write_out(XpB<E> &S, const XpBatchArgs &A, int k, int st)
    {
        const int tid = threadIdx.x;
        if (tid == 0 && A.maxv) ((double *)A.maxv)[k] = st == XP_SIX_SUCC ? S.tgtf[S.n] : 0.0;
        if (A.slack_sol)
            for (int j = tid; j < A.ldo; j += blockDim.x)
                ((double *)A.slack_sol)[(size_t)k * A.ldo + j] = j < S.C ? S.sol[j] : 0.0;
        if (A.tgtf_out)
            for (int j = tid; j < A.ldo; j += blockDim.x)
                ((double *)A.tgtf_out)[(size_t)k * A.ldo + j] = j < S.C ? S.tgtf[j] : 0.0;
    }
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 1024 / THREADS : 1)) k_batch_f64(XpBatchArgs A)
{
    xpb_kernel_body<OpsF64, false>(A);
}

// LPs beyond shared memory: same code, state slab in global memory
__global__ void __launch_bounds__(1024) k_batch_f64_gws(XpBatchArgs A)
{
    xpb_kernel_body<OpsF64, true>(A);
}

int pick_threads(int maxm, int maxn)
{
    if (const char *e = getenv("XP_BATCH_THREADS")) return atoi(e); // tuning knob
    long long cells = (long long)maxm * (maxn + maxm + 2);
    if (cells <= 16 * 64) return 64;
    if (cells <= 16 * 256) return 128;
    if (cells <= 16 * 1024) return 256;
    return 512;
}

// Register-resident fast path: one warp per LP (xp_batch_warp_f64.cuh).
template <int MR, int NS>
int launch_warp(xp_ctx *ctx, XpBatchArgs &A)
{
    int occ = 1;
    XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xpw::k_warp_f64<MR, NS>,
                                                                  32 * xpw::WARPS, 0));
    if (occ < 1) occ = 1;
    long long g = (long long)occ * ctx->sm_count;
    const long long need = ((long long)A.batch + xpw::WARPS - 1) / xpw::WARPS;
    if (g > need) g = need;
    XP_CUDA_OK(ctx, cudaMemsetAsync(A.queue, 0, sizeof(unsigned), ctx->stream));
    xpw::k_warp_f64<MR, NS><<<(unsigned)g, 32 * xpw::WARPS, 0, ctx->stream>>>(A);
    ctx->launches++;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

bool warp_path_enabled()
{
    const char *e = getenv("XP_BATCH_WARP"); // "0" forces the one-CTA-per-LP kernel (A/B tests)
    return !(e && e[0] == '0');
}

int launch_f64(xp_ctx *ctx, XpBatchArgs &A)
{
    if (A.maxm <= 32 && A.maxn + 1 + A.maxm <= 64 && warp_path_enabled()) {
        const bool one = A.maxn + 1 + A.maxm <= 32;
        if (A.maxm <= 8) return one ? launch_warp<8, 1>(ctx, A) : launch_warp<8, 2>(ctx, A);
        if (A.maxm <= 16) return one ? launch_warp<16, 1>(ctx, A) : launch_warp<16, 2>(ctx, A);
        if (A.maxm <= 24) return launch_warp<24, 2>(ctx, A);
        return launch_warp<32, 2>(ctx, A);
    }
    const size_t smem = xpb_smem_bytes(A.maxm, A.maxn, sizeof(double), sizeof(XpMinIdx));
    if (smem > ctx->smem_optin) {
        // Same kernel, state slab in global memory: one 1024-thread CTA per LP.
        const size_t stride = (smem + 255) & ~(size_t)255;
        long long g = 2LL * ctx->sm_count;
        if (g > A.batch) g = A.batch;
        void *ws = nullptr;
        int rc = xp_ctx_gws(ctx, stride * (size_t)g, &ws);
        if (rc) return rc;
        A.gws = (unsigned char *)ws;
        A.gws_stride = stride;
        XP_CUDA_OK(ctx, cudaMemsetAsync(A.queue, 0, sizeof(unsigned), ctx->stream));
        k_batch_f64_gws<<<(unsigned)g, 1024, 0, ctx->stream>>>(A);
        ctx->launches++;
        XP_CUDA_OK(ctx, cudaGetLastError());
        return 0;
    }
    XP_CUDA_OK(ctx, cudaMemsetAsync(A.queue, 0, sizeof(unsigned), ctx->stream));
    const int th = pick_threads(A.maxm, A.maxn);
    int occ = 1;
#define LAUNCH(TH)                                                                              \
    {                                                                                           \
        XP_CUDA_OK(ctx, cudaFuncSetAttribute(k_batch_f64<TH>,                                   \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                             (int)smem));                                       \
        XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_batch_f64<TH>, TH, \
                                                                      smem));                   \
        if (occ < 1) occ = 1;                                                                   \
        long long g = (long long)occ * ctx->sm_count;                                           \
        if (g > A.batch) g = A.batch;                                                           \
        k_batch_f64<TH><<<(unsigned)g, TH, smem, ctx->stream>>>(A);                             \
    }
    switch (th) {
    case 64: LAUNCH(64) break;
    case 128: LAUNCH(128) break;
    case 256: LAUNCH(256) break;
    default: LAUNCH(512) break;
    }
#undef LAUNCH
    ctx->launches++;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

} // namespace

// ---- device-pointer entry: inputs already resident in HBM ----
extern "C" int xp_six_two_stage_f64_batch_dev(xp_ctx *ctx, int batch, int m, int n,
                                              const double *d_leq, const double *d_tgtf,
                                              uint32_t max_iter, int rule, int32_t *d_status,
                                              double *d_maxv, double *d_slack_sol,
                                              double *d_tgtf_out, int32_t *d_eq2bv,
                                              uint32_t *d_iters, uint32_t *d_pivots)
{
    if (!ctx || batch < 0 || m < 1 || n < 1 || !d_leq || !d_tgtf) return XP_ERR_BAD_ARG;
    if (rule != XP_RULE_REFERENCE) return XP_ERR_BAD_ARG;
    if (batch == 0) return 0;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    void *scr = nullptr;
    int rc = xp_ctx_scratch(ctx, 256, &scr);
    if (rc) return rc;
    XpBatchArgs A;
    memset(&A, 0, sizeof A);
    A.batch = batch;
    A.m = m;
    A.n = n;
    A.leq = d_leq;
    A.tgtf = d_tgtf;
    A.max_iter = max_iter;
    A.ldo = n + m + 1;
    A.ldm = m;
    A.status = d_status;
    A.maxv = d_maxv;
    A.slack_sol = d_slack_sol;
    A.tgtf_out = d_tgtf_out;
    A.eq2bv = d_eq2bv;
    A.iters = d_iters;
    A.pivots = d_pivots;
    A.maxm = m;
    A.maxn = n;
    A.queue = (unsigned *)scr;
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    rc = launch_f64(ctx, A);
    if (rc) return rc;
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    XP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    XP_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev0, ctx->ev1));
    return 0;
}

// ---- host-pointer entry (uniform shape) ----
extern "C" int xp_six_two_stage_f64_batch(xp_ctx *ctx, int batch, int m, int n, const double *leq,
                                          const double *tgtf, uint32_t max_iter, int rule,
                                          int32_t *status, double *maxv, double *slack_sol,
                                          double *tgtf_out, int32_t *eq2bv, uint32_t *iters,
                                          uint32_t *pivots)
{
    if (!ctx || batch < 0 || m < 1 || n < 1 || !leq || !tgtf) return XP_ERR_BAD_ARG;
    if (rule != XP_RULE_REFERENCE) return XP_ERR_BAD_ARG;
    if (batch == 0) return 0;
    XpBatchHost H;
    H.batch = batch;
    H.m = m;
    H.n = n;
    H.leq = leq;
    H.tgtf = tgtf;
    H.leq_len = (size_t)batch * m * (n + 1);
    H.tgtf_len = (size_t)batch * (n + 1);
    H.max_iter = max_iter;
    H.ldo = n + m + 1;
    H.ldm = m;
    H.status = status;
    H.maxv = maxv;
    H.slack_sol = slack_sol;
    H.tgtf_out = tgtf_out;
    H.eq2bv = eq2bv;
    H.iters = iters;
    H.pivots = pivots;
    return xpb_host_run(ctx, H, launch_f64);
}

// ---- host-pointer entry (ragged shapes) ----
extern "C" int xp_six_two_stage_f64_ragged(xp_ctx *ctx, int batch, const int32_t *ms,
                                           const int32_t *ns, const int64_t *leq_off,
                                           const int64_t *tgtf_off, const double *leq,
                                           size_t leq_len, const double *tgtf, size_t tgtf_len,
                                           uint32_t max_iter, int rule, int ldo, int ldm,
                                           int32_t *status, double *maxv, double *slack_sol,
                                           double *tgtf_out, int32_t *eq2bv, uint32_t *iters,
                                           uint32_t *pivots)
{
    if (!ctx || batch < 0 || !ms || !ns || !leq_off || !tgtf_off || !leq || !tgtf)
        return XP_ERR_BAD_ARG;
    if (rule != XP_RULE_REFERENCE) return XP_ERR_BAD_ARG;
    if (batch == 0) return 0;
    XpBatchHost H;
    H.batch = batch;
    H.ms = ms;
    H.ns = ns;
    H.leq_off = leq_off;
    H.tgtf_off = tgtf_off;
    H.leq = leq;
    H.tgtf = tgtf;
    H.leq_len = leq_len;
    H.tgtf_len = tgtf_len;
    H.max_iter = max_iter;
    H.ldo = ldo;
    H.ldm = ldm;
    H.status = status;
    H.maxv = maxv;
    H.slack_sol = slack_sol;
    H.tgtf_out = tgtf_out;
    H.eq2bv = eq2bv;
    H.iters = iters;
    H.pivots = pivots;
    return xpb_host_run(ctx, H, launch_f64);
}
