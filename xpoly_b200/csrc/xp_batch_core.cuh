// One-CTA-per-LP simplex with the tableau in shared memory: the control
// skeleton of SIX<Mat,T>::TwoStageMethod shared by the FP64 kernel
// (xp_batch_f64.cu) and the exact fraction-free kernel (xp_batch_i64.cu).
//
// Reference (all /root/reference/src/com/lpsol.h): TwoStageMethod :1906-1930,
// stage1 :1783-1844, slack :1405-1433, constructBasicFeasibleSolution :838-988,
// solveSlackForm :1007-1191, findPivotBV :552-663, findPivotNVandBVPair
// :670-773, PivotPairTab :68-154.
//
// The arithmetic (element type, pivot update, ratio comparison, feasibility
// test, objective substitution) is supplied by an `Ops` policy.
#pragma once

#include "xp_common.cuh"

#include <cstring>

struct XpBatchArgs {
    int batch;
    int m, n;                         // uniform shape (when ms == nullptr)
    const int32_t *ms, *ns;           // ragged shapes (device) or nullptr
    const int64_t *leq_off, *tgtf_off; // ragged element offsets (device) or nullptr
    const void *leq, *tgtf;           // element pools (device)
    uint32_t max_iter;
    int ldo, ldm;                     // output strides (elements)
    int32_t *status;
    void *maxv, *slack_sol, *tgtf_out; // Ops-specific output layouts
    void *slack_sol2, *tgtf_out2;      // (denominators, exact path)
    int32_t *eq2bv;
    uint32_t *iters, *pivots;
    int maxm, maxn;                   // smem layout bounds
    unsigned *queue;                  // atomic work counter
    // LPs too large for shared memory: the same state block carved out of a
    // per-CTA slab of global memory instead (L2-resident for c5-sized tableaux).
    unsigned char *gws;
    size_t gws_stride;
};

constexpr int XPB_BIG = 0x7fffffff;

// Shared-memory resident solver state of one LP.
template <class E>
struct XpB {
    E *tab, *tgtf, *fcol, *sol;
    uint8_t *nvset;
    int *bv2eq, *eq2bv;
    uint32_t *tabu;
    int *row_cnt, *col_cnt;
    int *shi;    // 33 ints
    void *shk;   // 33 ratio keys
    int m, C, n, LD, W; // n = rhs_idx = C-1
    unsigned pivots;
    long long *misc;    // 8 spare 64-bit shared slots (exact path: D, overflow flag)
    int togk, togi;     // which half of shk / shi the next single-barrier reduction uses
};

__host__ __device__ inline size_t xpb_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bytes of dynamic shared memory for bounds (maxm, maxn), element size es,
// ratio-key size ks.  Layout mirrored in xpb_carve().
inline size_t xpb_smem_bytes(int maxm, int maxn, size_t es, size_t ks)
{
    const int nmax = maxn + 1 + maxm;   // variables incl. the auxiliary one
    const int LD = (nmax + 1) | 1;      // + constant column, odd => conflict-free column walks
    const int W = (nmax + 31) / 32;
    size_t b = 0;
    b += xpb_align((size_t)maxm * LD * es, 16);  // tab
    b += xpb_align((size_t)LD * es, 16) * 2;     // tgtf, sol
    b += xpb_align((size_t)maxm * es, 16);       // fcol
    b += xpb_align((size_t)33 * ks, 16);         // ratio keys
    b += xpb_align((size_t)(nmax + 1) * 4, 16);  // bv2eq
    b += xpb_align((size_t)maxm * 4, 16);        // eq2bv
    b += xpb_align((size_t)nmax * W * 4, 16);    // tabu
    b += xpb_align((size_t)nmax * 4, 16) * 2;    // row_cnt, col_cnt
    b += xpb_align(33 * 4, 16);                  // shi
    b += xpb_align((size_t)nmax + 1, 16);        // nvset
    b += 64;                                     // misc scalars
    return b;
}

#ifdef __CUDACC__

template <class E, class K>
__device__ inline void xpb_carve(XpB<E> &S, unsigned char *base, int maxm, int maxn, long long **misc)
{
    const int nmax = maxn + 1 + maxm;
    const int LD = (nmax + 1) | 1;
    const int W = (nmax + 31) / 32;
    size_t o = 0;
    S.tab = (E *)(base + o);
    o += xpb_align((size_t)maxm * LD * sizeof(E), 16);
    S.tgtf = (E *)(base + o);
    o += xpb_align((size_t)LD * sizeof(E), 16);
    S.sol = (E *)(base + o);
    o += xpb_align((size_t)LD * sizeof(E), 16);
    S.fcol = (E *)(base + o);
    o += xpb_align((size_t)maxm * sizeof(E), 16);
    S.shk = (void *)(base + o);
    o += xpb_align((size_t)33 * sizeof(K), 16);
    S.bv2eq = (int *)(base + o);
    o += xpb_align((size_t)(nmax + 1) * 4, 16);
    S.eq2bv = (int *)(base + o);
    o += xpb_align((size_t)maxm * 4, 16);
    S.tabu = (uint32_t *)(base + o);
    o += xpb_align((size_t)nmax * W * 4, 16);
    S.row_cnt = (int *)(base + o);
    o += xpb_align((size_t)nmax * 4, 16);
    S.col_cnt = (int *)(base + o);
    o += xpb_align((size_t)nmax * 4, 16);
    S.shi = (int *)(base + o);
    o += xpb_align(33 * 4, 16);
    S.nvset = (uint8_t *)(base + o);
    o += xpb_align((size_t)nmax + 1, 16);
    *misc = (long long *)(base + o);
    S.misc = *misc;
    S.LD = LD;
    S.W = W;
}

template <class E>
__device__ __forceinline__ bool xpb_tabu_get(const XpB<E> &S, int nv, int bv)
{
    return (S.tabu[nv * S.W + (bv >> 5)] >> (bv & 31)) & 1u;
}

// newPPT (lpsol.h:390-399): a fresh table per solveSlackForm call.
template <class E>
__device__ inline void xpb_tabu_reset(XpB<E> &S)
{
    for (int k = threadIdx.x; k < S.n * S.W; k += blockDim.x) S.tabu[k] = 0u;
    for (int k = threadIdx.x; k < S.n; k += blockDim.x) {
        S.row_cnt[k] = 0;
        S.col_cnt[k] = 0;
    }
    __syncthreads();
}

// PivotPairTab::disableNV (lpsol.h:114-121).
template <class E>
__device__ inline void xpb_disable_nv(XpB<E> &S, int q)
{
    const int n = S.n;
    for (int w = threadIdx.x; w < S.W; w += blockDim.x) {
        const int base = w << 5;
        uint32_t want = (n - base >= 32) ? 0xffffffffu : ((1u << (n - base)) - 1u);
        if ((q >> 5) == w) want &= ~(1u << (q & 31));
        uint32_t old = S.tabu[q * S.W + w];
        uint32_t add = want & ~old;
        S.tabu[q * S.W + w] = old | want;
        while (add) {
            int b = __ffs(add) - 1;
            add &= add - 1;
            S.col_cnt[base + b] += 1;
        }
    }
    if (threadIdx.x == 0) S.row_cnt[q] = n - 1;
    __syncthreads();
}

// Block reductions with ONE barrier: every warp leaves its result in one half of the scratch
// array, all threads combine the (<= 16) warp results themselves, and consecutive reductions
// alternate halves, so the barrier of the next one is what protects the slots from reuse.
// (CTAs of more than 512 threads -- the global-memory slab variant -- take the 3-barrier form.)

// Block arg-min over Ops::Key with the reference's "first strict minimum" rule.
template <class Ops>
__device__ inline typename Ops::Key xpb_block_best(XpB<typename Ops::E> &S, typename Ops::Key x)
{
    typedef typename Ops::Key K;
    K *sh = (K *)S.shk;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = Ops::better(x, Ops::shfl_xor(x, o));
    if (nw <= 16) {
        K *h = sh + (S.togk & 1) * 16;
        S.togk++;
        if (lane == 0) h[w] = x;
        __syncthreads();
        K y = h[0];
        for (int k = 1; k < nw; k++) y = Ops::better(y, h[k]);
        return y;
    }
    __syncthreads();
    if (lane == 0) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        K y = Ops::empty_key();
        if (lane < nw) y = sh[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) y = Ops::better(y, Ops::shfl_xor(y, o));
        if (lane == 0) sh[32] = y;
    }
    __syncthreads();
    return sh[32];
}

// Block min of an int and OR of a flag at once.
template <class E>
__device__ inline int xpb_min_int_or(XpB<E> &S, int x, int &flag)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    x = xp_warp_min_int(x);
    flag = __any_sync(0xffffffffu, flag);
    if (nw <= 8) {
        int *h = S.shi + (S.togi & 1) * 16;
        S.togi++;
        if (lane == 0) {
            h[w] = x;
            h[8 + w] = flag;
        }
        __syncthreads();
        int y = h[0], f = h[8];
        for (int k = 1; k < nw; k++) {
            y = min(y, h[k]);
            f |= h[8 + k];
        }
        flag = f;
        return y;
    }
    x = xp_block_min_int(x, S.shi);
    flag = __syncthreads_or(flag);
    return x;
}
template <class E>
__device__ inline int xpb_min_int(XpB<E> &S, int x)
{
    int f = 0;
    return xpb_min_int_or(S, x, f);
}

// findPivotBV (lpsol.h:552-663).  Returns the pivot ROW or -1.
template <class Ops>
__device__ inline int xpb_ratio_test(XpB<typename Ops::E> &S, int q)
{
    typedef typename Ops::Key K;
    const int n = S.n;
    K best = Ops::empty_key();
    for (int i = threadIdx.x; i < S.m; i += blockDim.x) { // pass 1, :571-612
        typename Ops::E a = S.tab[i * S.LD + q];
        if (Ops::le_zero(a)) continue;
        const int bv = S.eq2bv[i];
        if (xpb_tabu_get(S, q, bv)) continue;
        if (S.col_cnt[bv] >= n - 1) continue;
        best = Ops::better(best, Ops::make_key(S.tab[i * S.LD + n], a, i));
    }
    best = xpb_block_best<Ops>(S, best);
    if (Ops::key_index(best) >= 0) return Ops::key_index(best);
    best = Ops::empty_key();
    for (int i = threadIdx.x; i < S.m; i += blockDim.x) { // pass 2, :623-658
        const int bv = S.eq2bv[i];
        if (xpb_tabu_get(S, q, bv)) continue;
        if (S.col_cnt[bv] >= n - 1) continue;
        typename Ops::E a = S.tab[i * S.LD + q];
        if (Ops::is_zero(a)) continue;
        best = Ops::better(best, Ops::make_key(S.tab[i * S.LD + n], a, i));
    }
    best = xpb_block_best<Ops>(S, best);
    return Ops::key_index(best);
}

// SIX::pivot bookkeeping common to both arithmetics (:1504-1510).
template <class E>
__device__ __forceinline__ void xpb_swap_basis(XpB<E> &S, int p, int q, int bv)
{
    S.nvset[q] = 0;
    S.nvset[bv] = 1;
    S.eq2bv[p] = q;
    S.bv2eq[q] = p;
    S.bv2eq[bv] = -1;
}

// solveSlackForm (lpsol.h:1007-1191).  Returns the SIX status; *iters = cnt.
template <class Ops>
__device__ inline int xpb_solve_loop(XpB<typename Ops::E> &S, uint32_t max_iter, uint32_t *iters)
{
    typedef typename Ops::E E;
    const int tid = threadIdx.x;
    xpb_tabu_reset(S);
    const int n = S.n;
    for (int j = tid; j < S.C; j += blockDim.x) S.sol[j] = Ops::zero(); // sol.reinit, :1028
    uint32_t cnt = 0;
    while (cnt < max_iter) {
        int q = -1, p = -1;
        for (;;) {
            // pricing, :1054-1069
            int best = XPB_BIG, anypos = 0;
            for (int j = tid; j < n; j += blockDim.x) {
                if (S.nvset[j] && Ops::pos(S.tgtf[j])) {
                    anypos = 1;
                    if (best == XPB_BIG && S.row_cnt[j] < n - 1) best = j;
                }
            }
            best = xpb_min_int_or(S, best, anypos);
            const int zlim = best == XPB_BIG ? n : best;
            for (int j = tid; j < zlim; j += blockDim.x) // only basic entries change and nobody reads
                if (!S.nvset[j]) S.tgtf[j] = Ops::zero(); // those before the barriers of pivot(), :1059
            if (best == XPB_BIG) {
                if (!anypos) { // optimal exit, :1089-1127
                    *iters = cnt;
                    return Ops::optimal_exit(S);
                }
                // findPivotNVandBVPair, :670-773 (pass B skips the c_j > 0 columns
                // that already failed in pass A: findPivotBV is pure)
                int found = 0;
                for (int pass = 0; pass < 2 && !found; pass++) {
                    int last = -1;
                    for (;;) {
                        int cand = XPB_BIG;
                        for (int j = last + 1 + tid; j < n; j += blockDim.x) {
                            if (!S.nvset[j] || S.row_cnt[j] >= n - 1) continue;
                            E c = S.tgtf[j];
                            bool take = pass == 0 ? Ops::pos(c) : (!Ops::pos(c) && Ops::is_zero(c));
                            if (take) {
                                cand = j;
                                break;
                            }
                        }
                        cand = xpb_min_int(S, cand);
                        if (cand == XPB_BIG) break;
                        int r = xpb_ratio_test<Ops>(S, cand);
                        if (r >= 0) {
                            q = cand;
                            p = r;
                            found = 1;
                            break;
                        }
                        last = cand;
                    }
                }
                if (!found) {
                    *iters = cnt;
                    return XP_SIX_UNBOUND; // :1138-1141
                }
                break;
            }
            q = best;
            p = xpb_ratio_test<Ops>(S, q);
            if (p >= 0) break;
            xpb_disable_nv(S, q); // :1146-1151
        }
        const int bv = S.eq2bv[p]; // (the ratio test's barrier is behind every read of the table)
        if (tid == 0) { // genPair, :1156
            uint32_t *w = &S.tabu[q * S.W + (bv >> 5)];
            const uint32_t bit = 1u << (bv & 31);
            if (!(*w & bit)) {
                *w |= bit;
                S.row_cnt[q] += 1;
                S.col_cnt[bv] += 1;
            }
        }
        int rc = Ops::pivot(S, p, q); // :1170
        if (rc) {
            *iters = cnt;
            return rc;
        }
        cnt++;
    }
    *iters = cnt;
    return XP_SIX_TIME_OUT;
}

// TwoStageMethod for LP `k` of the batch.  Returns the status.
template <class Ops>
__device__ inline int xpb_two_stage(XpB<typename Ops::E> &S, const XpBatchArgs &A,
                                    const typename Ops::In *leq, const typename Ops::In *tg,
                                    int m, int n, uint32_t *iters)
{
    typedef typename Ops::E E;
    const int tid = threadIdx.x;
    const int LD = S.LD;
    // stage1 decision, :1794-1803
    int pos = 0, bneg = 0;
    for (int j = tid; j < n; j += blockDim.x) pos |= Ops::in_pos(tg[j]);
    for (int i = tid; i < m; i += blockDim.x) bneg |= Ops::in_neg(leq[(size_t)i * (n + 1) + n]);
    pos = __syncthreads_or(pos);
    bneg = __syncthreads_or(bneg);
    const bool aux = !pos || bneg;
    const int xa = n;                   // auxiliary variable index (if any)
    const int s0 = aux ? n + 1 : n;     // first slack column
    S.m = m;
    S.n = s0 + m;
    S.C = S.n + 1;
    S.pivots = 0;
    Ops::reset(S);
    // slack form [A | (-1) | I | b], :860-875 / :1405-1433, and the identity basis
    for (int e = tid; e < m * S.C; e += blockDim.x) {
        const int i = e / S.C, j = e - i * S.C;
        E v;
        if (j < n) v = Ops::from_in(leq[(size_t)i * (n + 1) + j]);
        else if (aux && j == xa) v = Ops::from_int(-1);
        else if (j < S.n) v = Ops::from_int(j - s0 == i ? 1 : 0);
        else v = Ops::from_in(leq[(size_t)i * (n + 1) + n]);
        S.tab[i * LD + j] = v;
    }
    for (int j = tid; j < S.C; j += blockDim.x) {
        E v = Ops::zero();
        if (aux) {
            if (j == xa) v = Ops::from_int(-1);
        } else if (j < n) v = Ops::from_in(tg[j]);
        else if (j == S.n) v = Ops::from_in(tg[n]);
        S.tgtf[j] = v;
        if (j < S.n) {
            S.nvset[j] = j < s0;
            S.bv2eq[j] = j < s0 ? -1 : j - s0;
        }
    }
    for (int i = tid; i < m; i += blockDim.x) S.eq2bv[i] = s0 + i;
    __syncthreads();

    if (aux) {
        // forced first pivot on the row of the first minimum constant term, :892-908
        int prow = Ops::argmin_rhs(S);
        int rc = Ops::pivot(S, prow, xa);
        if (rc) return rc;
        uint32_t it1 = 0;
        int st = xpb_solve_loop<Ops>(S, A.max_iter, &it1);
        if (st < 0) return st;
        if (st != XP_SIX_SUCC) return XP_SIX_NO_PRI_FEASIBLE_SOL; // :912-915
        if (!Ops::is_zero(S.tgtf[S.n])) return XP_SIX_NO_PRI_FEASIBLE_SOL; // :919-922
        __syncthreads();
        if (!S.nvset[xa]) { // xa still basic: pivot it out, :924-941
            const int eqnum = S.bv2eq[xa];
            int cand = XPB_BIG;
            for (int j = tid; j < S.n; j += blockDim.x)
                if (S.nvset[j] && !Ops::is_zero(S.tab[eqnum * LD + j])) {
                    cand = j;
                    break;
                }
            cand = xpb_min_int(S, cand);
            if (cand == XPB_BIG) return XP_ERR_REFERENCE_UB; // reference ASSERTs (:937)
            rc = Ops::pivot(S, eqnum, cand);
            if (rc) return rc;
        }
        // restore the original objective by substitution (:944-953), then drop
        // column xa and re-index the maps (:956-986)
        rc = Ops::restore_objective(S, tg, n);
        if (rc) return rc;
        __syncthreads();
        {
            // shift columns (xa, C) one to the left; one warp per row, ascending
            // 32-wide chunks (read chunk, then write it one slot lower)
            const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
            for (int i = w; i < m; i += nw) {
                for (int j0 = xa + 1; j0 < S.C; j0 += 32) {
                    const int j = j0 + lane;
                    E t = Ops::zero();
                    if (j < S.C) t = S.tab[i * LD + j];
                    __syncwarp();
                    if (j < S.C) S.tab[i * LD + j - 1] = t;
                    __syncwarp();
                }
            }
            if (w == 0) {
                for (int j0 = xa + 1; j0 < S.C; j0 += 32) {
                    const int j = j0 + lane;
                    E t = Ops::zero();
                    uint8_t nv = 0;
                    int b2e = 0;
                    if (j < S.C) t = S.tgtf[j];
                    if (j < S.n) {
                        nv = S.nvset[j];
                        b2e = S.bv2eq[j];
                    }
                    __syncwarp();
                    if (j < S.C) S.tgtf[j - 1] = t;
                    if (j < S.n) {
                        S.nvset[j - 1] = nv;
                        S.bv2eq[j - 1] = b2e;
                    }
                    __syncwarp();
                }
            }
        }
        for (int i = tid; i < m; i += blockDim.x)
            if (S.eq2bv[i] > xa) S.eq2bv[i] -= 1;
        S.n -= 1;
        S.C -= 1;
        __syncthreads();
    }
    return xpb_solve_loop<Ops>(S, A.max_iter, iters);
}

// Persistent CTAs pull LP indices from an atomic queue, so long-running LPs
// (unbounded ones exit only after exhausting the tabu table) do not hold up
// the rest of the grid.
// GWS = false keeps every state pointer in the shared address space at compile time
// (LDS/STS instead of generic loads); GWS = true is the global-memory slab variant.
template <class Ops, bool GWS>
__device__ inline void xpb_kernel_body(const XpBatchArgs &A)
{
    typedef typename Ops::E E;
    extern __shared__ __align__(16) unsigned char xpb_smem[];
    __shared__ int s_lp;
    XpB<E> S;
    long long *misc;
    unsigned char *state = GWS ? A.gws + (size_t)blockIdx.x * A.gws_stride : xpb_smem;
    xpb_carve<E, typename Ops::Key>(S, state, A.maxm, A.maxn, &misc);
    Ops::bind(S, misc);
    S.togk = S.togi = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_lp = (int)atomicAdd(A.queue, 1u);
        __syncthreads();
        const int k = s_lp;
        if (k >= A.batch) break;
        const int m = A.ms ? A.ms[k] : A.m;
        const int n = A.ns ? A.ns[k] : A.n;
        const typename Ops::In *leq = (const typename Ops::In *)A.leq +
                                      (A.leq_off ? A.leq_off[k] : (int64_t)k * m * (n + 1));
        const typename Ops::In *tg = (const typename Ops::In *)A.tgtf +
                                     (A.tgtf_off ? A.tgtf_off[k] : (int64_t)k * (n + 1));
        uint32_t iters = 0;
        int st = xpb_two_stage<Ops>(S, A, leq, tg, m, n, &iters);
        __syncthreads();
        Ops::write_out(S, A, k, st);
        if (threadIdx.x == 0) {
            if (A.status) A.status[k] = st;
            if (A.iters) A.iters[k] = iters;
            if (A.pivots) A.pivots[k] = S.pivots;
        }
        if (A.eq2bv)
            for (int i = threadIdx.x; i < m; i += blockDim.x) A.eq2bv[(size_t)k * A.ldm + i] = S.eq2bv[i];
    }
}

#endif // __CUDACC__

// ---------------------------------------------------------------------------
// Host plumbing shared by the FP64 and exact entry points: stage inputs in the
// ctx scratch block, launch, copy results back.  All pools hold 8-byte elements.
// ---------------------------------------------------------------------------
struct XpBatchHost {
    int batch = 0;
    int m = 0, n = 0;                       // uniform shape, or
    const int32_t *ms = nullptr, *ns = nullptr; // ragged shapes (host)
    const int64_t *leq_off = nullptr, *tgtf_off = nullptr;
    const void *leq = nullptr, *tgtf = nullptr; // host pools
    size_t leq_len = 0, tgtf_len = 0;       // elements
    uint32_t max_iter = 0;
    int ldo = 0, ldm = 0;
    int maxv_elems = 1;                     // 8-byte elements per LP in maxv
    // host outputs (any may be null)
    int32_t *status = nullptr;
    void *maxv = nullptr, *slack_sol = nullptr, *slack_sol2 = nullptr;
    void *tgtf_out = nullptr, *tgtf_out2 = nullptr;
    int32_t *eq2bv = nullptr;
    uint32_t *iters = nullptr, *pivots = nullptr;
};

typedef int (*XpBatchLaunch)(xp_ctx *ctx, XpBatchArgs &A);

inline int xpb_host_run(xp_ctx *ctx, const XpBatchHost &H, XpBatchLaunch launch)
{
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t B = (size_t)H.batch;
    int maxm = H.m, maxn = H.n;
    if (H.ms) {
        maxm = maxn = 0;
        for (int k = 0; k < H.batch; k++) {
            if (H.ms[k] < 1 || H.ns[k] < 1) return XP_ERR_BAD_ARG;
            maxm = H.ms[k] > maxm ? H.ms[k] : maxm;
            maxn = H.ns[k] > maxn ? H.ns[k] : maxn;
            if (H.ms[k] + H.ns[k] + 1 > H.ldo) return XP_ERR_BAD_ARG;
            // every LP must lie inside the pools (offsets and lengths in elements)
            if (!H.leq_off || !H.tgtf_off || H.leq_off[k] < 0 || H.tgtf_off[k] < 0) return XP_ERR_BAD_ARG;
            if ((size_t)H.leq_off[k] + (size_t)H.ms[k] * (H.ns[k] + 1) > H.leq_len) return XP_ERR_BAD_ARG;
            if ((size_t)H.tgtf_off[k] + (size_t)H.ns[k] + 1 > H.tgtf_len) return XP_ERR_BAD_ARG;
        }
        if (H.ldm < maxm) return XP_ERR_BAD_ARG;
    }
    XP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t ldo = (size_t)H.ldo, ldm = (size_t)H.ldm;
    size_t total = 1024 + pad(H.leq_len * 8) + pad(H.tgtf_len * 8) + 4 * pad(B * ldo * 8) +
                   pad(B * 8 * H.maxv_elems) + pad(B * ldm * 4) + 5 * pad(B * 4) +
                   2 * pad(B * 8) + 8192;
    void *scr = nullptr;
    int rc = xp_ctx_scratch(ctx, total, &scr);
    if (rc) return rc;
    unsigned char *base = (unsigned char *)scr;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        off = (off + 255) & ~(size_t)255;
        void *p = base + off;
        off += bytes;
        return p;
    };
    unsigned *queue = (unsigned *)take(256);
    void *d_leq = take(H.leq_len * 8), *d_tg = take(H.tgtf_len * 8);
    int32_t *d_ms = nullptr, *d_ns = nullptr;
    int64_t *d_lo = nullptr, *d_to = nullptr;
    cudaStream_t s = ctx->stream;
    if (H.ms) {
        d_ms = (int32_t *)take(B * 4);
        d_ns = (int32_t *)take(B * 4);
        d_lo = (int64_t *)take(B * 8);
        d_to = (int64_t *)take(B * 8);
        XP_CUDA_OK(ctx, cudaMemcpyAsync(d_ms, H.ms, B * 4, cudaMemcpyHostToDevice, s));
        XP_CUDA_OK(ctx, cudaMemcpyAsync(d_ns, H.ns, B * 4, cudaMemcpyHostToDevice, s));
        XP_CUDA_OK(ctx, cudaMemcpyAsync(d_lo, H.leq_off, B * 8, cudaMemcpyHostToDevice, s));
        XP_CUDA_OK(ctx, cudaMemcpyAsync(d_to, H.tgtf_off, B * 8, cudaMemcpyHostToDevice, s));
    }
    void *d_sol = H.slack_sol ? take(B * ldo * 8) : nullptr;
    void *d_sol2 = H.slack_sol2 ? take(B * ldo * 8) : nullptr;
    void *d_tgo = H.tgtf_out ? take(B * ldo * 8) : nullptr;
    void *d_tgo2 = H.tgtf_out2 ? take(B * ldo * 8) : nullptr;
    void *d_maxv = H.maxv ? take(B * 8 * H.maxv_elems) : nullptr;
    int32_t *d_e2b = H.eq2bv ? (int32_t *)take(B * ldm * 4) : nullptr;
    int32_t *d_st = H.status ? (int32_t *)take(B * 4) : nullptr;
    uint32_t *d_it = H.iters ? (uint32_t *)take(B * 4) : nullptr;
    uint32_t *d_pv = H.pivots ? (uint32_t *)take(B * 4) : nullptr;
    // Large uniform batches of small LPs are pipelined: the pool is cut into chunks, chunk c+1
    // is uploaded (copy stream) while chunk c is being solved, and every chunk runs on its own
    // stream so that the long tail of one chunk (unbounded LPs exit only after exhausting the
    // tabu table) overlaps the next chunk's work.  Small or ragged batches: one chunk on the
    // ctx stream.
    const size_t in_bytes = (H.leq_len + H.tgtf_len) * 8;
    int nc = 1;
    if (!H.ms && in_bytes >= ((size_t)48 << 20) && (size_t)maxm * (maxn + maxm + 2) * 8 <= (64u << 10)) {
        nc = (int)(in_bytes >> 26) + 1; // ~64 MB per chunk
        if (nc > XP_PIPE_MAX) nc = XP_PIPE_MAX;
        if ((size_t)nc > B / 1024) nc = (int)(B / 1024);
        if (nc < 1) nc = 1;
    }
    if (nc > 1) {
        rc = xp_ctx_pipe(ctx);
        if (rc) return rc;
        XP_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_begin, s));
        XP_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->pipe_copy, ctx->pipe_begin, 0));
    }
    const size_t per_leq = nc > 1 ? (size_t)H.m * (H.n + 1) : 0, per_tg = nc > 1 ? (size_t)H.n + 1 : 0;
    for (int c = 0; c < nc; c++) {
        const size_t k0 = B * c / nc, k1 = B * (c + 1) / nc, nb = k1 - k0;
        cudaStream_t cs = nc > 1 ? ctx->pipe_stream[c] : s; // solve + download of this chunk
        cudaStream_t up = nc > 1 ? ctx->pipe_copy : s;      // uploads, in chunk order
        if (nc > 1) {
            XP_CUDA_OK(ctx, cudaMemcpyAsync((char *)d_leq + k0 * per_leq * 8, (const char *)H.leq + k0 * per_leq * 8,
                                            nb * per_leq * 8, cudaMemcpyHostToDevice, up));
            XP_CUDA_OK(ctx, cudaMemcpyAsync((char *)d_tg + k0 * per_tg * 8, (const char *)H.tgtf + k0 * per_tg * 8,
                                            nb * per_tg * 8, cudaMemcpyHostToDevice, up));
            XP_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_up[c], up));
            XP_CUDA_OK(ctx, cudaStreamWaitEvent(cs, ctx->pipe_begin, 0));
            XP_CUDA_OK(ctx, cudaStreamWaitEvent(cs, ctx->pipe_up[c], 0));
        } else {
            XP_CUDA_OK(ctx, cudaMemcpyAsync(d_leq, H.leq, H.leq_len * 8, cudaMemcpyHostToDevice, s));
            XP_CUDA_OK(ctx, cudaMemcpyAsync(d_tg, H.tgtf, H.tgtf_len * 8, cudaMemcpyHostToDevice, s));
        }
        XpBatchArgs A;
        memset(&A, 0, sizeof A);
        A.batch = (int)nb;
        A.m = H.m;
        A.n = H.n;
        A.ms = d_ms;
        A.ns = d_ns;
        A.leq_off = d_lo;
        A.tgtf_off = d_to;
        A.leq = (char *)d_leq + k0 * per_leq * 8;
        A.tgtf = (char *)d_tg + k0 * per_tg * 8;
        A.max_iter = H.max_iter;
        A.ldo = H.ldo;
        A.ldm = H.ldm;
#define XPB_AT(ptr, stride) ((ptr) ? (void *)((char *)(ptr) + k0 * (size_t)(stride)) : nullptr)
        A.status = (int32_t *)XPB_AT(d_st, 4);
        A.maxv = XPB_AT(d_maxv, 8 * H.maxv_elems);
        A.slack_sol = XPB_AT(d_sol, ldo * 8);
        A.slack_sol2 = XPB_AT(d_sol2, ldo * 8);
        A.tgtf_out = XPB_AT(d_tgo, ldo * 8);
        A.tgtf_out2 = XPB_AT(d_tgo2, ldo * 8);
        A.eq2bv = (int32_t *)XPB_AT(d_e2b, ldm * 4);
        A.iters = (uint32_t *)XPB_AT(d_it, 4);
        A.pivots = (uint32_t *)XPB_AT(d_pv, 4);
        A.maxm = maxm;
        A.maxn = maxn;
        A.queue = queue + c; // one work-queue counter per chunk (chunks overlap)
        if (c == 0) XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, cs));
        ctx->stream = cs; // the launcher issues on ctx->stream
        rc = launch(ctx, A);
        ctx->stream = s;
        if (rc) return rc;
        if (nc == 1) XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, s));
#undef XPB_AT
    }
    // Downloads are issued only after every chunk is queued: a D2H copy into pageable memory
    // blocks the calling thread until it has completed.
    for (int c = 0; c < nc; c++) {
        const size_t k0 = B * c / nc, k1 = B * (c + 1) / nc, nb = k1 - k0;
        cudaStream_t cs = nc > 1 ? ctx->pipe_stream[c] : s;
#define XPB_AT(ptr, stride) ((ptr) ? (void *)((char *)(ptr) + k0 * (size_t)(stride)) : nullptr)
#define XPB_D2H(dst, src, stride)                                                               \
    if (dst)                                                                                    \
    XP_CUDA_OK(ctx, cudaMemcpyAsync((char *)(dst) + k0 * (size_t)(stride), XPB_AT(src, stride), \
                                    nb * (size_t)(stride), cudaMemcpyDeviceToHost, cs))
        XPB_D2H(H.status, d_st, 4);
        XPB_D2H(H.maxv, d_maxv, 8 * H.maxv_elems);
        XPB_D2H(H.slack_sol, d_sol, ldo * 8);
        XPB_D2H(H.slack_sol2, d_sol2, ldo * 8);
        XPB_D2H(H.tgtf_out, d_tgo, ldo * 8);
        XPB_D2H(H.tgtf_out2, d_tgo2, ldo * 8);
        XPB_D2H(H.eq2bv, d_e2b, ldm * 4);
        XPB_D2H(H.iters, d_it, 4);
        XPB_D2H(H.pivots, d_pv, 4);
#undef XPB_D2H
#undef XPB_AT
        if (nc > 1) {
            XP_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_done[c], cs));
            XP_CUDA_OK(ctx, cudaStreamWaitEvent(s, ctx->pipe_done[c], 0));
        }
    }
    // pipelined: ev0 .. ev1 spans first kernel start .. last chunk complete
    if (nc > 1) XP_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, s));
    XP_CUDA_OK(ctx, cudaStreamSynchronize(s));
    XP_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev0, ctx->ev1));
    return 0;
}
