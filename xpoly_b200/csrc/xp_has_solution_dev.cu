// Lineq::has_solution (linsys.cpp:830-906) entirely on the device: one WARP per dependence
// query runs both MIPs of the query (max, then min if max did not succeed) including the
// depth-first branch & bound of MIP::RecusivePart (lpsol.h:2426-2612), every node relaxation
// solved by the register-resident exact TwoStageMethod of xp_batch_warp_i64.cuh.
//
// Why: these systems are tiny (2-6 variables, 3-10 rows).  The lock-step path of xp_entry.cu
// needs ~10 host round trips per batch (one per B&B wave) and spends >95 % of its time
// normalising / scaling / unpacking on the host -- more CPU time per query than the reference
// needs to solve it.  Here a query is one unit of work: the host converts the batch to int64
// once, launches once, and reads one int32 per query.
//
// What runs here for a query (leq m x (n+1) integers, vc = -I | 0, no equalities):
//   * the all-ones objective over the columns that occur (linsys.cpp:852-857, reviseTargetFunc);
//   * MIP maxm / minm = DFS branch & bound replayed in the reference's order (fork_count and the
//     incumbent are order dependent, :2474-2497): explicit stack, a node's LP = the query's rows
//     + one branching row per stack level (:2504-2565);
//   * SIX::maxm = TwoStageMethod on the node LP + calcFinalSolution (:1850-1899);
//     SIX::minm = TwoStageMethod on the explicit dual (calcDualMaxm :1585-1655) + the primal
//     solution read out of the final dual objective row (:1713-1716);
//   * is_satisfying (:2363-2408, integrality of every entry), bound and fork tests (:2474-2497).
// Values are exact rationals num / D (D = common denominator of the LP that produced them);
// comparisons cross-multiply in 128 bits -- the value semantics of the reference's Rational
// wherever that one stays exact.  An int64 overflow anywhere reports XP_ERR_OVERFLOW for the query.
// Queries that do not fit (equalities, rational inputs, more than 24 rows / 32 columns at the
// deepest node) are answered by the lock-step path of xp_entry.cu.  Two instantiations: 16 rows
// (237 registers) and 24 rows (255 registers, a few spilled words) for the larger dependence
// polyhedra (13 rows x 4 variables after the reference's pre-filter, tests/golden/deppoly_queries.json).
#include "xp_batch_warp_i64.cuh"

#include <vector>

namespace {

using namespace xpwi;

constexpr int HS_MRMAX = 24; // rows of the deepest node LP (and variables of its dual): kernels for 16 and 24
constexpr int HS_N1 = 16;    // n + 1 at most
constexpr int HS_DEPTH = 18;

struct HsArgs {
    int batch;
    const int32_t *ns, *ms;
    const int64_t *off; // into pool, in int64 elements
    const i64 *pool;
    int is_int, is_unique;
    int32_t *result;
    unsigned *queue;
};

struct HsFrame {
    int stage, col, sol_ceil, have_tmp;
    i64 tv_num, tv_den; // value of the floor child (tmpv, :2527-2543)
};

template <int HS_MR>
struct HsWarp {
    i64 leq[HS_MR * HS_N1];        // node LP: the query's rows, then one branching row per level
    i64 dual[HS_N1 * (HS_MR + 1)]; // explicit dual of the node LP (min problems): n rows x (m + 1)
    i64 tg[HS_N1], dtg[HS_MR + 1];
    i64 sol[HS_N1];                // numerators of the node's final solution over D
    HsFrame fr[HS_DEPTH];
    int fork[HS_N1 + 1];
};

__device__ __forceinline__ bool q_lt(i64 an, i64 ad, i64 bn, i64 bd) { return (i128)an * bd < (i128)bn * ad; }
__device__ __forceinline__ bool q_le(i64 an, i64 ad, i64 bn, i64 bd) { return (i128)an * bd <= (i128)bn * ad; }

// One node relaxation: SIX::maxm / minm on rows [0, m) of S.leq.  Returns the SIX status; on
// SIX_SUCC S.sol[0..n] holds the solution numerators over D and (v_num, D) the objective value.
template <int HS_MR>
__device__ int solve_node(HsWarp<HS_MR> &S, LP<HS_MR, 1> &W, const XpBatchArgs &A, bool is_max, int m, int n, i64 &v_num,
                          i64 &D, bool &ovf)
{
    const int lane = W.lane;
    uint32_t it = 0;
    int st;
    if (is_max) {
        st = two_stage(W, A, S.leq, S.tg, m, n, &it);
        if (st != XP_SIX_SUCC) return st;
        D = (i64)W.dv.D;
        if (lane < n) S.sol[lane] = W.sol0; // slack_sol of the structural columns (calcFinalSolution, :1880-1884)
    } else {
        // dual: one variable per row, one row per variable: [-A^T | c], objective -b (:1602-1623)
        for (int e = lane; e < n * (m + 1); e += 32) {
            const int i = e / (m + 1), j = e - i * (m + 1);
            S.dual[e] = j < m ? -S.leq[j * (n + 1) + i] : S.tg[i];
        }
        if (lane <= m) S.dtg[lane] = lane < m ? -S.leq[lane * (n + 1) + n] : 0;
        __syncwarp();
        st = two_stage(W, A, S.dual, S.dtg, n, m, &it);
        if (st != XP_SIX_SUCC) return st;
        D = (i64)W.dv.D;
        // y_k = -(coefficient of dual slack k in the final dual objective row), :1713-1716
        const i64 c = W.c0;
        const int src = m + (lane < n ? lane : 0);
        const i64 ck = __shfl_sync(FULL, c, src & 31);
        if (lane < n) S.sol[lane] = -ck;
    }
    if (lane == 0) S.sol[n] = D; // the constant entry of the solution is 1 (:1885-1887)
    __syncwarp();
    // v = sum_j sol_j * tgtf_j over the caller's objective, constant column included (:1896-1898)
    i128 acc = 0;
    for (int j = 0; j <= n; j++) acc += (i128)S.sol[j] * S.tg[j];
    if (acc > (i128)0x7fffffffffffffffLL || acc < -(i128)0x7fffffffffffffffLL) ovf = true;
    v_num = (i64)acc;
    return XP_SIX_SUCC;
}

// MIP::maxm / minm (RecusivePart, lpsol.h:2426-2612) on the query in S.leq[0..m0).  Returns the IP status.
template <int HS_MR>
__device__ int run_mip(HsWarp<HS_MR> &S, LP<HS_MR, 1> &W, const XpBatchArgs &A, bool is_max, bool is_int, int m0, int n)
{
    const int lane = W.lane;
    if (lane <= HS_N1) S.fork[lane] = 0;
    if (lane == 0) S.fr[0].stage = 0;
    __syncwarp();
    int depth = 1;
    bool has_best = false, ovf = false;
    i64 best_n = 0, best_d = 1, v_n = 0, v_d = 1;
    for (;;) {
        // ---- the frame on top is at stage 0: solve its LP and consume the result (feed) ----
        const int m = m0 + depth - 1;
        i64 vn = 0, D = 1;
        const int six = solve_node(S, W, A, is_max, m, n, vn, D, ovf);
        int status;
        bool descend = false;
        if (ovf) {
            status = XP_ERR_OVERFLOW;
        } else if (six < 0) {
            status = six;
        } else if (six != XP_SIX_SUCC) { // :2450-2466
            v_n = 0, v_d = 1;
            status = six == XP_SIX_UNBOUND ? XP_IP_UNBOUND : XP_IP_NO_PRI_FEASIBLE_SOL;
        } else {
            v_n = vn, v_d = D;
            int col = -1; // is_satisfying: the first entry that is not an integer (:2363-2408)
            if (is_int) {
                const bool frac = lane <= n && (S.sol[lane] % D) != 0;
                const unsigned bad = __ballot_sync(FULL, frac);
                col = bad ? __ffs(bad) - 1 : -1;
            }
            if (col < 0) {
                status = XP_IP_SUCC;
            } else if (has_best && (is_max ? q_le(v_n, v_d, best_n, best_d) : q_le(best_n, best_d, v_n, v_d))) {
                status = XP_IP_NO_BETTER_THAN_BEST_SOL; // :2474-2485
            } else if (S.fork[col] >= 1) {
                status = XP_IP_NO_PRI_FEASIBLE_SOL; // :2486-2496
            } else {
                if (depth + 1 >= HS_DEPTH || m + 1 > HS_MR) {
                    status = XP_ERR_TOO_LARGE; // cannot happen for pre-screened queries (one fork per variable)
                } else {
                    const int fl = (int)(S.sol[col] / D); // typecast2int truncates
                    __syncwarp();
                    if (lane == 0) {
                        S.fork[col] += 1;
                        S.fr[depth - 1].col = col;
                        S.fr[depth - 1].sol_ceil = fl + 1;
                        S.fr[depth - 1].stage = 1;
                        S.fr[depth - 1].have_tmp = 0;
                        S.fr[depth].stage = 0;
                    }
                    if (lane <= n) S.leq[m * (n + 1) + lane] = lane == col ? 1 : (lane == n ? fl : 0); // x_col <= floor, :2514-2520
                    __syncwarp();
                    depth++;
                    descend = true;
                    status = 0;
                }
            }
        }
        if (descend) continue;
        // ---- finish(status): unwind until a frame has another child to run, :2527-2611 ----
        for (;;) {
            depth--;
            if (depth == 0) return status;
            const HsFrame f = S.fr[depth - 1];
            if (f.stage == 1) { // the floor child just returned
                if (status < 0) continue;
                __syncwarp();
                if (lane == 0) {
                    if (status == XP_IP_SUCC) {
                        S.fr[depth - 1].tv_num = v_n;
                        S.fr[depth - 1].tv_den = v_d;
                        S.fr[depth - 1].have_tmp = 1;
                    }
                    S.fr[depth - 1].stage = 2;
                    S.fr[depth].stage = 0;
                }
                if (status == XP_IP_SUCC && (!has_best || (is_max ? q_lt(best_n, best_d, v_n, v_d) : q_lt(v_n, v_d, best_n, best_d))))
                    has_best = true, best_n = v_n, best_d = v_d;
                const int mrow = m0 + depth - 1; // the ceil child always runs: -x_col <= -ceil, :2555-2563
                if (lane <= n) S.leq[mrow * (n + 1) + lane] = lane == f.col ? -1 : (lane == n ? -f.sol_ceil : 0);
                __syncwarp();
                depth++;
                break;
            }
            // stage 2: the ceil child returned, :2563-2611
            if (status < 0) continue;
            if (status == XP_IP_SUCC) {
                if (f.have_tmp && (is_max ? q_lt(v_n, v_d, f.tv_num, f.tv_den) : q_lt(f.tv_num, f.tv_den, v_n, v_d)))
                    v_n = f.tv_num, v_d = f.tv_den;
                if (!has_best || (is_max ? q_lt(best_n, best_d, v_n, v_d) : q_lt(v_n, v_d, best_n, best_d)))
                    has_best = true, best_n = v_n, best_d = v_d;
            } else if (f.have_tmp) {
                v_n = f.tv_num, v_d = f.tv_den;
                if (!has_best || (is_max ? q_lt(best_n, best_d, v_n, v_d) : q_lt(v_n, v_d, best_n, best_d)))
                    has_best = true, best_n = v_n, best_d = v_d;
                status = XP_IP_SUCC;
            }
        }
    }
}

template <int HS_MR>
__global__ void __launch_bounds__(32 * WARPS) k_has_solution(HsArgs H)
{
    __shared__ __align__(16) i64 scratch[WARPS][32];
    __shared__ HsWarp<HS_MR> state[WARPS];
    LP<HS_MR, 1> W;
    W.c1 = W.sol1 = 0;
    W.t1 = 0ull;
    W.rc1 = W.cc1 = 0;
    W.b2e1 = -1;
    W.lane = threadIdx.x & 31;
    W.sc = scratch[threadIdx.x >> 5];
    HsWarp<HS_MR> &S = state[threadIdx.x >> 5];
    XpBatchArgs A;
    A.max_iter = 10000u; // six.set_param(m_indent, 10000), :2441
    const int lane = W.lane;
    for (;;) {
        int k = 0;
        if (lane == 0) k = (int)atomicAdd(H.queue, 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= H.batch) break;
        const int m0 = H.ms[k], n = H.ns[k];
        const i64 *src = H.pool + H.off[k];
        __syncwarp();
        for (int e = lane; e < m0 * (n + 1); e += 32) S.leq[e] = src[e];
        __syncwarp();
        if (lane <= n) { // objective: 1 for every variable that occurs, 0 for the constant (linsys.cpp:852-857)
            i64 nz = 0;
            if (lane < n)
                for (int i = 0; i < m0; i++) nz |= S.leq[i * (n + 1) + lane];
            S.tg[lane] = nz != 0 ? 1 : 0;
        }
        __syncwarp();
        int res = 0;
        for (int pass = 0; pass < 2; pass++) { // max first, then min (:864-881)
            const int st = run_mip(S, W, A, pass == 0, H.is_int != 0, m0, n);
            if (st < 0) {
                res = st;
                break;
            }
            if (st == 0 || (!H.is_unique && st == 1)) { // IP_SUCC == SIX_SUCC == 0, IP_UNBOUND == SIX_UNBOUND == 1
                res = 1;
                break;
            }
        }
        if (lane == 0) H.result[k] = res;
    }
}

} // namespace

// Host side: `sel` lists the queries of the batch that qualify (integer entries, no equalities,
// small enough); their rows are packed as int64 and answered by one launch per kernel size.
template <int HS_MR>
static int hs_launch(xp_ctx *ctx, const HsArgs &H, int B)
{
    int occ = 1;
    XP_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_has_solution<HS_MR>, 32 * WARPS, 0));
    if (occ < 1) occ = 1;
    long long g = (long long)occ * ctx->sm_count;
    const long long need = ((long long)B + WARPS - 1) / WARPS;
    if (g > need) g = need;
    k_has_solution<HS_MR><<<(unsigned)g, 32 * WARPS, 0, ctx->stream>>>(H);
    ctx->launches++;
    return 0;
}

int xp_has_solution_device(xp_ctx *ctx, const std::vector<int> &sel_in, const int32_t *ns, const int32_t *ms,
                           const int64_t *leq_off, const xp_rat *leq_pool, int is_int_sol, int is_unique_sol,
                           int32_t *result)
{
    if (sel_in.empty()) return 0;
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    // queries whose deepest node has at most 16 rows first, the 17..24-row ones after them
    std::vector<int> sel;
    sel.reserve(sel_in.size());
    for (int b : sel_in)
        if (ms[b] + ns[b] <= 16) sel.push_back(b);
    const int B16 = (int)sel.size();
    for (int b : sel_in)
        if (ms[b] + ns[b] > 16) sel.push_back(b);
    const int B = (int)sel.size();
    std::vector<int32_t> hn(B), hm(B);
    std::vector<int64_t> off(B);
    size_t total = 0;
    for (int s = 0; s < B; s++) {
        hn[s] = ns[sel[s]];
        hm[s] = ms[sel[s]];
        off[s] = (int64_t)total;
        total += (size_t)hm[s] * (hn[s] + 1);
    }
    std::vector<int64_t> pool(total);
    for (int s = 0; s < B; s++) {
        const xp_rat *p = leq_pool + leq_off[sel[s]];
        int64_t *d = pool.data() + off[s];
        const size_t cnt = (size_t)hm[s] * (hn[s] + 1);
        for (size_t e = 0; e < cnt; e++) d[e] = p[e].num; // den == 1 (pre-screened)
    }
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t bytes = pad(total * 8) + 2 * pad((size_t)B * 4) + pad((size_t)B * 8) + pad((size_t)B * 4) + 1024;
    void *scr = nullptr;
    int rc = xp_ctx_scratch(ctx, bytes, &scr);
    if (rc) return rc;
    unsigned char *base = (unsigned char *)scr;
    size_t o = 0;
    auto take = [&](size_t b) {
        void *p = base + o;
        o += pad(b);
        return p;
    };
    i64 *d_pool = (i64 *)take(total * 8);
    int32_t *d_n = (int32_t *)take((size_t)B * 4), *d_m = (int32_t *)take((size_t)B * 4);
    int64_t *d_off = (int64_t *)take((size_t)B * 8);
    int32_t *d_res = (int32_t *)take((size_t)B * 4);
    unsigned *d_q = (unsigned *)take(512);
    cudaStream_t s = ctx->stream;
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_pool, pool.data(), total * 8, cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_n, hn.data(), (size_t)B * 4, cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_m, hm.data(), (size_t)B * 4, cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemcpyAsync(d_off, off.data(), (size_t)B * 8, cudaMemcpyHostToDevice, s));
    XP_CUDA_OK(ctx, cudaMemsetAsync(d_q, 0, 512, s));
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, s));
    HsArgs H;
    H.pool = d_pool;
    H.is_int = is_int_sol;
    H.is_unique = is_unique_sol;
    if (B16 > 0) {
        H.batch = B16;
        H.ns = d_n, H.ms = d_m, H.off = d_off, H.result = d_res, H.queue = d_q;
        rc = hs_launch<16>(ctx, H, B16);
        if (rc) return rc;
    }
    if (B > B16) {
        H.batch = B - B16;
        H.ns = d_n + B16, H.ms = d_m + B16, H.off = d_off + B16, H.result = d_res + B16, H.queue = d_q + 64;
        rc = hs_launch<HS_MRMAX>(ctx, H, B - B16);
        if (rc) return rc;
    }
    XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, s));
    XP_CUDA_OK(ctx, cudaGetLastError());
    std::vector<int32_t> hres(B);
    XP_CUDA_OK(ctx, cudaMemcpyAsync(hres.data(), d_res, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
    XP_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev0, ctx->ev1));
    for (int k = 0; k < B; k++) result[sel[k]] = hres[k];
    return 0;
}

// true if query b can take the device path
bool xp_has_solution_device_fits(int n, int m, int k, const xp_rat *leq)
{
    if (k != 0 || n < 1 || m < 1) return false;
    // deepest node: one branching row per variable (fork_count allows each variable once)
    const int rows = m + n;
    if (n + 1 > HS_N1 || rows > HS_MRMAX) return false;
    if (n + rows + 2 > 32) return false; // columns incl. slacks, auxiliary variable and constant (primal and dual alike)
    const size_t cnt = (size_t)m * (n + 1);
    for (size_t e = 0; e < cnt; e++)
        if (leq[e].den != 1) return false;
    return true;
}
