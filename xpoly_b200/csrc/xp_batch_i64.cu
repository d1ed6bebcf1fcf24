// Batched exact simplex, one CTA per LP: the fraction-free integer twin of
// SIX<RMat,Rational>::TwoStageMethod.
//
// The reference's rational tableau entry is a_ij = N_ij / D with one common
// denominator D > 0 per LP (D = 1 for integer input); N is int64, every product
// is formed in 128 bits, and the division by the previous D is exact
// (integer-preserving / Bareiss-Edmonds pivoting):
//     pivot on (p,q), P = N_pq, s = sign(P):
//       i != p :  N_ij <- (N_ij*|P| - s*N_iq*N_pj) / D
//       row p  :  N_pj <- s*N_pj
//       c_j    :  Nt_j <- (Nt_j*|P| - s*Nt_q*N_pj) / D          (j <  rhs)
//       const  :  Nt_r <- (Nt_r*|P| + s*Nt_q*N_pr) / D          (lpsol.h:1496-1501)
//       D <- |P|
// All decisions (pricing sign tests, cross-multiplied ratio comparisons,
// tolerant-free equality) are the value semantics of the reference's Rational
// (rational.cpp:229-397) whenever it did not fall into its lossy appro() path.
// Results are reduced to canonical num/den on output.  Entries leaving int64
// are detected (XP_ERR_OVERFLOW), never wrapped.
//
// Exact division uses the odd-part modular inverse of D (one Newton iteration
// chain per pivot, by one thread), so the per-element cost is two 64x64->128
// products, a 128-bit subtract, a funnel shift and one 64-bit multiply.
#include "xp_batch_core.cuh"
#include "xp_exact_arith.cuh"
#include "xp_batch_warp_i64.cuh"

#include <cstdlib>

namespace {

typedef long long i64;
typedef unsigned long long u64;
typedef __int128 i128;

struct KeyI64 {
    i64 num, den; // ratio num/den with den > 0
    int i;
};

// misc slots (shared): [0] D, [1] overflow flag, [2] inverse of D's odd part, [3] tz(D)
enum { M_D = 0, M_OVF = 1, M_INV = 2, M_TZ = 3 };

using xpx::ff_div;
using xpx::gcd64;
using xpx::inv_odd64;

struct OpsI64 {
    typedef i64 E;
    typedef i64 In;
    typedef KeyI64 Key;

    __device__ static __forceinline__ E zero() { return 0; }
    __device__ static __forceinline__ E from_int(int i) { return (i64)i; }
    __device__ static __forceinline__ E from_in(In x) { return x; }
    __device__ static __forceinline__ bool in_pos(In x) { return x > 0; }
    __device__ static __forceinline__ bool in_neg(In x) { return x < 0; }
    __device__ static __forceinline__ bool pos(E x) { return x > 0; }     // D > 0
    __device__ static __forceinline__ bool le_zero(E x) { return x <= 0; }
    __device__ static __forceinline__ bool is_zero(E x) { return x == 0; }

    __device__ static __forceinline__ Key empty_key()
    {
        Key k;
        k.num = 0;
        k.den = 1;
        k.i = -1;
        return k;
    }
    __device__ static __forceinline__ Key make_key(E b, E a, int i)
    { // v = rhs / coeff as a sign-normalised fraction (common D cancels)
        Key k;
        k.num = a < 0 ? -b : b;
        k.den = a < 0 ? -a : a;
        k.i = i;
        return k;
    }
    __device__ static __forceinline__ Key better(Key a, Key b)
    { // first strict minimum: smaller value, lower index on ties (lpsol.h:603-611)
        if (b.i < 0) return a;
        if (a.i < 0) return b;
        const i128 l = (i128)b.num * a.den, r = (i128)a.num * b.den; // b.v < a.v ?
        if (l < r || (l == r && b.i < a.i)) return b;
        return a;
    }
    __device__ static __forceinline__ Key shfl_xor(Key x, int o)
    {
        Key y;
        y.num = __shfl_xor_sync(0xffffffffu, x.num, o);
        y.den = __shfl_xor_sync(0xffffffffu, x.den, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        return y;
    }
    __device__ static __forceinline__ int key_index(Key k) { return k.i; }

    __device__ static __forceinline__ void bind(XpB<E> &, long long *) {}
    __device__ static void reset(XpB<E> &S)
    {
        __syncthreads();
        if (threadIdx.x == 0) {
            S.misc[M_D] = 1;
            S.misc[M_OVF] = 0;
            S.misc[M_INV] = 1;
            S.misc[M_TZ] = 0;
        }
        __syncthreads();
    }

    __device__ static int argmin_rhs(XpB<E> &S)
    {
        Key best = empty_key();
        for (int i = threadIdx.x; i < S.m; i += blockDim.x) {
            Key k;
            k.num = S.tab[i * S.LD + S.n];
            k.den = 1;
            k.i = i;
            best = better(best, k);
        }
        best = xpb_block_best<OpsI64>(S, best);
        return best.i;
    }

    __device__ static int pivot(XpB<E> &S, int p, int q)
    {
        const int tid = threadIdx.x, LD = S.LD, C = S.C, n = S.n, m = S.m;
        const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
        const int bv = S.eq2bv[p];
        const i64 P = S.tab[p * LD + q];
        const i64 cq = S.tgtf[q];
        xpx::Div dv;
        dv.set((u64)S.misc[M_D], (u64)S.misc[M_INV], (int)S.misc[M_TZ]);
        const i64 aP = P < 0 ? -P : P;
        const bool sneg = P < 0;
        __syncthreads();
        for (int i = tid; i < m; i += blockDim.x)
            if (i != p) {
                i64 f = S.tab[i * LD + q];
                S.fcol[i] = sneg ? -f : f; // s * N_iq
            }
        __syncthreads();
        bool ovf = false;
        const i64 *rowp = S.tab + p * LD;
        const i64 scq = sneg ? -cq : cq;
        for (int j = tid; j < C; j += blockDim.x) { // objective row
            i128 x = (i128)S.tgtf[j] * aP;
            i128 y = (i128)scq * rowp[j];
            x = j >= n ? x + y : x - y;
            S.tgtf[j] = ff_div(x, dv, ovf);
        }
        for (int i = w; i < m; i += nw) { // integer-preserving elimination
            if (i == p) continue;
            const i64 f = S.fcol[i];
            i64 *row = S.tab + i * LD;
            for (int j = lane; j < C; j += 32) {
                i128 x = (i128)row[j] * aP - (i128)f * rowp[j];
                row[j] = ff_div(x, dv, ovf);
            }
        }
        __syncthreads();
        if (sneg)
            for (int j = tid; j < C; j += blockDim.x) S.tab[p * LD + j] = -S.tab[p * LD + j];
        if (tid == 0) {
            xpb_swap_basis(S, p, q, bv);
            S.misc[M_D] = aP;
            const int t = __ffsll(aP) - 1;
            S.misc[M_TZ] = t;
            S.misc[M_INV] = (i64)inv_odd64((u64)aP >> t);
        }
        S.pivots++;
        return __syncthreads_or(ovf ? 1 : 0) ? XP_ERR_OVERFLOW : 0;
    }

    // Optimal exit.  In exact arithmetic the row-sum half of is_feasible
    // (lpsol.h:805-814) is an identity (basic columns are unit vectors), so the
    // answer is decided by the sign test on the basic values (:798-802).
    __device__ static int optimal_exit(XpB<E> &S)
    {
        const int tid = threadIdx.x, LD = S.LD, n = S.n;
        int bad = 0;
        for (int j = tid; j < S.C; j += blockDim.x) {
            i64 v = (j < n && !S.nvset[j]) ? S.tab[S.bv2eq[j] * LD + n] : 0;
            S.sol[j] = v;
            if (v < 0) bad = 1;
        }
        bad = __syncthreads_or(bad);
        return bad ? XP_SIX_OPTIMAL_IS_INFEASIBLE : XP_SIX_SUCC;
    }

    // lpsol.h:944-953 in exact arithmetic: each basic variable i with c_i != 0
    // contributes -c_i * row(i) (+ on the constant column); basic columns are
    // unit vectors, so the substitutions commute and c_i is the input value.
    __device__ static int restore_objective(XpB<E> &S, const In *tg, int n_orig)
    {
        const int tid = threadIdx.x, LD = S.LD, C = S.C, rhs = S.n;
        const i64 D = S.misc[M_D];
        int ovf = 0;
        __syncthreads();
        for (int j = tid; j < C; j += blockDim.x) {
            i128 acc = 0;
            if (j < n_orig) acc = (i128)tg[j] * D;
            else if (j == rhs) acc = (i128)tg[n_orig] * D;
            for (int i = 0; i < n_orig; i++) {
                const i64 ci = tg[i];
                if (ci == 0 || S.nvset[i]) continue;
                const i128 t = (i128)ci * S.tab[S.bv2eq[i] * LD + j];
                acc = j >= rhs ? acc + t : acc - t;
            }
            if (acc > (i128)0x7fffffffffffffffLL || acc < -(i128)0x7fffffffffffffffLL) ovf = 1;
            S.sol[j] = (i64)acc; // staged: tgtf is still being read by nobody, but keep it simple
        }
        __syncthreads();
        for (int j = tid; j < C; j += blockDim.x) S.tgtf[j] = S.sol[j];
        ovf = __syncthreads_or(ovf);
        return ovf ? XP_ERR_OVERFLOW : 0;
    }

    __device__ static void write_out(XpB<E> &S, const XpBatchArgs &A, int k, int st)
    {
        const int tid = threadIdx.x;
        const i64 D = S.misc[M_D];
        if (tid == 0 && A.maxv) {
            i64 num = st == XP_SIX_SUCC ? S.tgtf[S.n] : 0, den = 1;
            if (num != 0) {
                i64 g = gcd64(num, D);
                num /= g;
                den = D / g;
            }
            ((i64 *)A.maxv)[2 * (size_t)k] = num;
            ((i64 *)A.maxv)[2 * (size_t)k + 1] = den;
        }
        for (int j = tid; j < A.ldo; j += blockDim.x) {
            const size_t o = (size_t)k * A.ldo + j;
            if (A.slack_sol) {
                i64 num = j < S.C ? S.sol[j] : 0, den = 1;
                if (num != 0) {
                    i64 g = gcd64(num, D);
                    num /= g;
                    den = D / g;
                }
                ((i64 *)A.slack_sol)[o] = num;
                if (A.slack_sol2) ((i64 *)A.slack_sol2)[o] = den;
            }
            if (A.tgtf_out) {
                i64 num = j < S.C ? S.tgtf[j] : 0, den = 1;
                if (num != 0) {
                    i64 g = gcd64(num, D);
                    num /= g;
                    den = D / g;
                }
                ((i64 *)A.tgtf_out)[o] = num;
                if (A.tgtf_out2) ((i64 *)A.tgtf_out2)[o] = den;
            }
        }
    }
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 1024 / THREADS : 1)) k_batch_i64(XpBatchArgs A)
{
    xpb_kernel_body<OpsI64, false>(A);
}

// LPs beyond shared memory: same code, state slab in global memory
__global__ void __launch_bounds__(1024) k_batch_i64_gws(XpBatchArgs A)
{
    xpb_kernel_body<OpsI64, true>(A);
}

int pick_threads_i64(int maxm, int maxn)
{
    if (const char *e = getenv("XP_BATCH_THREADS_I64")) return atoi(e); // tuning knob
    long long cells = (long long)maxm * (maxn + maxm + 2);
    if (cells <= 16 * 64) return 64;
    if (cells <= 16 * 256) return 128;
    return 256;
}

// Register-resident fast path: one warp per LP (xp_batch_warp_i64.cuh).
template <int MR, int NS>
int launch_warp_i64(xp_ctx *ctx, XpBatchArgs &A)
{
    int occ = 1;
    XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xpwi::k_warp_i64<MR, NS>,
                                                                  32 * xpwi::WARPS, 0));
    if (occ < 1) occ = 1;
    long long g = (long long)occ * ctx->sm_count;
    const long long need = ((long long)A.batch + xpwi::WARPS - 1) / xpwi::WARPS;
    if (g > need) g = need;
    XP_CUDA_OK(ctx, cudaMemsetAsync(A.queue, 0, sizeof(unsigned), ctx->stream));
    xpwi::k_warp_i64<MR, NS><<<(unsigned)g, 32 * xpwi::WARPS, 0, ctx->stream>>>(A);
    ctx->launches++;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

int launch_i64(xp_ctx *ctx, XpBatchArgs &A)
{
    // Measured on B200: with 128-bit products the register kernel only pays while a lane's share
    // of the tableau leaves room for four warps per scheduler -- dependence-feasibility systems
    // (~7 x 11: 1.4x the CTA kernel); at 24 x 48 it runs at 255 registers with spills and loses
    // (100 000 LPs: 126 ms against 72 ms), so larger shapes stay on the one-CTA-per-LP kernel.
    const char *we = getenv("XP_BATCH_WARP"); // "0" forces the CTA kernel, "2" the warp kernel (A/B tests)
    const bool fits = A.maxm <= 32 && A.maxn + 1 + A.maxm <= 64;
    const bool one = A.maxn + 1 + A.maxm <= 32;
    const bool pays = A.maxm <= 8 || (A.maxm <= 16 && one);
    if (fits && !(we && we[0] == '0') && (pays || (we && we[0] == '2'))) {
        if (A.maxm <= 8) return one ? launch_warp_i64<8, 1>(ctx, A) : launch_warp_i64<8, 2>(ctx, A);
        if (A.maxm <= 16) return one ? launch_warp_i64<16, 1>(ctx, A) : launch_warp_i64<16, 2>(ctx, A);
        if (A.maxm <= 24) return launch_warp_i64<24, 2>(ctx, A);
        return launch_warp_i64<32, 2>(ctx, A);
    }
    const size_t smem = xpb_smem_bytes(A.maxm, A.maxn, sizeof(i64), sizeof(KeyI64));
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
        k_batch_i64_gws<<<(unsigned)g, 1024, 0, ctx->stream>>>(A);
        ctx->launches++;
        XP_CUDA_OK(ctx, cudaGetLastError());
        return 0;
    }
    XP_CUDA_OK(ctx, cudaMemsetAsync(A.queue, 0, sizeof(unsigned), ctx->stream));
    const int th = pick_threads_i64(A.maxm, A.maxn);
    int occ = 1;
#define LAUNCH(TH)                                                                              \
    {                                                                                           \
        XP_CUDA_OK(ctx, cudaFuncSetAttribute(k_batch_i64<TH>,                                   \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                             (int)smem));                                       \
        XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_batch_i64<TH>, TH, \
                                                                      smem));                   \
        if (occ < 1) occ = 1;                                                                   \
        long long g = (long long)occ * ctx->sm_count;                                           \
        if (g > A.batch) g = A.batch;                                                           \
        k_batch_i64<TH><<<(unsigned)g, TH, smem, ctx->stream>>>(A);                             \
    }
    switch (th) {
    case 32: LAUNCH(32) break; // (measured at c4: 15.5 ms against 7.8 ms with 128 threads -- the pivot is work-bound)
    case 64: LAUNCH(64) break;
    case 128: LAUNCH(128) break;
    default: LAUNCH(256) break;
    }
#undef LAUNCH
    ctx->launches++;
    XP_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

} // namespace

extern "C" int xp_six_two_stage_i64_batch(xp_ctx *ctx, int batch, int m, int n, const int64_t *leq,
                                          const int64_t *tgtf, uint32_t max_iter, int rule,
                                          int32_t *status, int64_t *maxv_num_den,
                                          int64_t *slack_sol_num, int64_t *slack_sol_den,
                                          int64_t *tgtf_out_num, int64_t *tgtf_out_den,
                                          int32_t *eq2bv, uint32_t *iters, uint32_t *pivots)
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
    H.maxv_elems = 2;
    H.status = status;
    H.maxv = maxv_num_den;
    H.slack_sol = slack_sol_num;
    H.slack_sol2 = slack_sol_den;
    H.tgtf_out = tgtf_out_num;
    H.tgtf_out2 = tgtf_out_den;
    H.eq2bv = eq2bv;
    H.iters = iters;
    H.pivots = pivots;
    return xpb_host_run(ctx, H, launch_i64);
}

extern "C" int xp_six_two_stage_i64_ragged(xp_ctx *ctx, int batch, const int32_t *ms,
                                           const int32_t *ns, const int64_t *leq_off,
                                           const int64_t *tgtf_off, const int64_t *leq,
                                           size_t leq_len, const int64_t *tgtf, size_t tgtf_len,
                                           uint32_t max_iter, int rule, int ldo, int ldm,
                                           int32_t *status, int64_t *maxv_num_den,
                                           int64_t *slack_sol_num, int64_t *slack_sol_den,
                                           int64_t *tgtf_out_num, int64_t *tgtf_out_den,
                                           int32_t *eq2bv, uint32_t *iters, uint32_t *pivots)
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
    H.maxv_elems = 2;
    H.status = status;
    H.maxv = maxv_num_den;
    H.slack_sol = slack_sol_num;
    H.slack_sol2 = slack_sol_den;
    H.tgtf_out = tgtf_out_num;
    H.tgtf_out2 = tgtf_out_den;
    H.eq2bv = eq2bv;
    H.iters = iters;
    H.pivots = pivots;
    return xpb_host_run(ctx, H, launch_i64);
}
