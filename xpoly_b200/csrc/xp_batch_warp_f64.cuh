// One-WARP-per-LP FP64 simplex with the tableau in REGISTERS: the fast path of
// SIX<FloatMat,Float>::TwoStageMethod for LPs of at most 32 rows and 64 variables
// (structural + auxiliary + slack), i.e. the config-2 shape (tableau 32 x 64) and
// everything smaller.  Same control skeleton, same arithmetic and therefore the
// same bits as the one-CTA-per-LP kernel (xp_batch_core.cuh + OpsF64); what
// changes is where the state lives and that no block barrier is left:
//
//   * lane l owns tableau columns l and 32+l ("slots" 0 and 1): a[i][s] for all
//     rows i, objective coefficients c[s], tabu rows / counters / bv2eq of those
//     variables -- 2 x MR doubles of tableau per lane, indexed statically;
//   * the constant column and eq2bv are kept one row per lane (b, e2b), which
//     is what the ratio test wants;
//   * the entering column is turned from "one lane holds it" into "one row per
//     lane" through a 256-byte per-warp scratch line (MR stores by the owner,
//     one load per lane), and the rank-1 update reads its multipliers from the
//     same line as broadcast loads;
//   * the pivot row of a lane's own columns is picked out of the register file
//     with a warp-uniform switch (p is uniform), and written back the same way;
//   * pricing, the tabu table (PivotPairTab, lpsol.h:68-154) and the basis maps
//     are ballots / 64-bit masks.
//
// The bound is the FP64 pipe: 2 x (MR+1) x 64 non-fused mul/add per pivot.
//
// Reference (all /root/reference/src/com/lpsol.h): TwoStageMethod :1906-1930,
// stage1 :1783-1844, slack :1405-1433, constructBasicFeasibleSolution :838-988,
// solveSlackForm :1007-1191, findPivotBV :552-663, findPivotNVandBVPair
// :670-773, pivot :1455-1511, is_feasible :783-822, PivotPairTab :68-154.
#pragma once

#include "xp_batch_core.cuh"

#ifdef __CUDACC__

namespace xpw {

typedef unsigned long long u64;
constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS = 4; // warps per CTA (each warp is independent)

__device__ __forceinline__ u64 shfl_u64(u64 x, int src)
{
    unsigned lo = __shfl_sync(FULL, (unsigned)x, src), hi = __shfl_sync(FULL, (unsigned)(x >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 ballot64(bool p0, bool p1)
{
    return (u64)__ballot_sync(FULL, p0) | ((u64)__ballot_sync(FULL, p1) << 32);
}
__device__ __forceinline__ u64 low_mask(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

// Float's tolerant predicates (flty.cpp:41-95) against the constants 0 and 1, reduced to one
// comparison each; case by case the same truth value as xp_feq / xp_fle, NaN included:
//   xp_feq(a, 0) <=> |a| <= eps        xp_fle(a, 0) <=> a <= eps
//   xp_feq(a, 1) <=> |a - 1| <= eps    (a - 1 rounded once, as flty.cpp:57-58 does)
__device__ __forceinline__ bool is_zero(double a) { return fabs(a) <= XP_EPS; }
__device__ __forceinline__ bool le_zero(double a) { return a <= XP_EPS; }
__device__ __forceinline__ bool is_one(double a) { return fabs(__dadd_rn(a, -1.0)) <= XP_EPS; }

// Warp arg-min of (v, lane) over the lanes with `on` set, ties to the lowest lane (the
// reference's first strict minimum, lpsol.h:608): the doubles are mapped to order-preserving
// 64-bit integers (-0 folded into +0 first) and reduced with two REDUX steps instead of a
// five-level shuffle butterfly.  Returns -1 when no lane is on.
__device__ __forceinline__ int warp_argmin(double v, bool on)
{
    const long long bits = __double_as_longlong(__dadd_rn(v, 0.0));
    int hi = (int)(bits >> 32);
    unsigned lo = (unsigned)bits;
    if (hi < 0) {
        hi ^= 0x7fffffff;
        lo = ~lo;
    }
    if (!on) hi = 0x7fffffff;
    const int mh = __reduce_min_sync(FULL, hi);
    const bool c1 = on && hi == mh;
    const unsigned ml = __reduce_min_sync(FULL, c1 ? lo : 0xffffffffu);
    const unsigned win = __ballot_sync(FULL, c1 && lo == ml);
    return win ? __ffs((int)win) - 1 : -1;
}

#define XPW_ROWS(X)                                                                            \
    X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) \
    X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31)

// Per-lane state.  Everything a lane owns per column slot is a separate scalar (x0 for column
// `lane`, x1 for column 32+lane) and every access is written out per slot: any loop or
// comparison chain over a small member array is turned into a dynamically indexed access by
// the compiler, which would push the whole structure into local memory.
template <int MR, int NS>
struct LP {
    double a0[MR], a1[MR]; // tableau(i, lane), tableau(i, 32 + lane)
    double c0, c1;         // objective row, own columns
    double sol0, sol1;     // slack solution, own columns (zero until an optimal exit)
    double cr;             // objective constant tgtf[rhs] (uniform)
    double b;              // constant column, row `lane`
    int e2b;               // eq2bv[lane]
    int b2e0, b2e1;        // bv2eq of own columns
    u64 t0, t1;            // tabu rows of own columns (bit bv)
    int rc0, rc1, cc0, cc1; // PivotPairTab row / column population of own columns
    u64 nvm;               // non-basic variables (uniform)
    u64 rowfull, colfull;  // row_cnt[j] >= n-1 / col_cnt[j] >= n-1 (uniform)
    int m, n;              // rows, variables (= rhs_idx)
    unsigned pivots;
    int lane;
    double *sc;            // this warp's scratch line (32 doubles, shared memory)
    static constexpr int kNS = NS;
};

// M(s) for slot 0 and, with two slots, slot 1
#define XPW_EACH(M) \
    M(0)            \
    if (NS == 2) { M(1) }

template <int NS, class T>
__device__ __forceinline__ T pick(T x0, T x1, int slot)
{
    if (NS == 1) return x0;
    return slot ? x1 : x0;
}
// value of a per-column quantity at (uniform) column j, broadcast to all lanes
template <int NS>
__device__ __forceinline__ double at_col(double x0, double x1, int j)
{
    return __shfl_sync(FULL, pick<NS>(x0, x1, j >> 5), j & 31);
}
template <int NS>
__device__ __forceinline__ int at_col_i(int x0, int x1, int j)
{
    return __shfl_sync(FULL, pick<NS>(x0, x1, j >> 5), j & 31);
}
template <int NS>
__device__ __forceinline__ u64 colmask(bool p0, bool p1)
{
    return NS == 1 ? (u64)__ballot_sync(FULL, p0) : ballot64(p0, p1);
}

// Column q, held by lane q&31, into the scratch line: sc[i] = a(i, q).
template <int MR, int NS>
__device__ __forceinline__ void col_to_scratch(LP<MR, NS> &W, int q)
{
    __syncwarp();
    if (W.lane == (q & 31)) {
        if (NS == 1 || q < 32) {
#pragma unroll
            for (int i = 0; i < MR; i += 2)
                *reinterpret_cast<double2 *>(W.sc + i) = make_double2(W.a0[i], W.a0[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < MR; i += 2)
                *reinterpret_cast<double2 *>(W.sc + i) = make_double2(W.a1[i], W.a1[i + 1]);
        }
    }
    __syncwarp();
}

// Row p (uniform) of a lane's own columns out of / into the register file: a jump table of
// register moves.  The moves are opaque (asm) so that the switch cannot be folded back into
// an indexed load.
__device__ __forceinline__ double opaque(double x)
{
    double y;
    asm volatile("mov.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
template <int MR, int NS>
__device__ __forceinline__ void get_row(const LP<MR, NS> &W, int p, double &r0, double &r1)
{
    r0 = 0.0;
    r1 = 0.0;
    switch (p) {
#define XPW_GET(i)                              \
    case i:                                     \
        if constexpr (i < MR) {                 \
            r0 = opaque(W.a0[i]);               \
            if (NS == 2) r1 = opaque(W.a1[i]);  \
        }                                       \
        break;
        XPW_ROWS(XPW_GET)
#undef XPW_GET
    default: break;
    }
}
template <int MR, int NS>
__device__ __forceinline__ void set_row(LP<MR, NS> &W, int p, double r0, double r1)
{
    switch (p) {
#define XPW_SET(i)                              \
    case i:                                     \
        if constexpr (i < MR) {                 \
            W.a0[i] = opaque(r0);               \
            if (NS == 2) W.a1[i] = opaque(r1);  \
        }                                       \
        break;
        XPW_ROWS(XPW_SET)
#undef XPW_SET
    default: break;
    }
}

// newPPT (lpsol.h:390-399)
template <class L>
__device__ __forceinline__ void tabu_reset(L &W)
{
    W.t0 = W.t1 = 0ull;
    W.rc0 = W.rc1 = W.cc0 = W.cc1 = 0;
    const u64 all = (0 >= W.n - 1) ? low_mask(W.n) : 0ull;
    W.rowfull = all;
    W.colfull = all;
}
template <class L>
__device__ __forceinline__ void tabu_masks(L &W)
{
    constexpr int NS = L::kNS;
    const int lim = W.n - 1, j0 = W.lane, j1 = 32 + W.lane;
    W.rowfull = colmask<NS>(j0 < W.n && W.rc0 >= lim, j1 < W.n && W.rc1 >= lim);
    W.colfull = colmask<NS>(j0 < W.n && W.cc0 >= lim, j1 < W.n && W.cc1 >= lim);
}
// PivotPairTab::genPair (lpsol.h:100), counters kept as in the CTA kernel.  The counters only
// grow between resets, so the two "full" masks are updated in place: one shuffle carries
// (newly set, row now full) from the owner of q, one ballot the column flag of bv's owner.
template <class L>
__device__ __forceinline__ void tabu_gen_pair(L &W, int q, int bv)
{
    constexpr int NS = L::kNS;
    const int lim = W.n - 1;
    int info = 0; // bit 0: pair newly set, bit 1: row q full now
    if (W.lane == (q & 31)) {
        const u64 bit = 1ull << bv;
        if (NS == 1 || q < 32) {
            if (!(W.t0 & bit)) {
                W.t0 |= bit;
                W.rc0 += 1;
                info = 1 | (W.rc0 >= lim ? 2 : 0);
            }
        } else {
            if (!(W.t1 & bit)) {
                W.t1 |= bit;
                W.rc1 += 1;
                info = 1 | (W.rc1 >= lim ? 2 : 0);
            }
        }
    }
    info = __shfl_sync(FULL, info, q & 31);
    if (info) {
        bool cfull = false;
        if (W.lane == (bv & 31)) {
            if (NS == 1 || bv < 32) {
                W.cc0 += 1;
                cfull = W.cc0 >= lim;
            } else {
                W.cc1 += 1;
                cfull = W.cc1 >= lim;
            }
        }
        if (info & 2) W.rowfull |= 1ull << q;
        if (__any_sync(FULL, cfull)) W.colfull |= 1ull << bv;
    }
}
// PivotPairTab::disableNV (lpsol.h:114-121)
template <class L>
__device__ __forceinline__ void tabu_disable_nv(L &W, int q)
{
    constexpr int NS = L::kNS;
    const u64 want = low_mask(W.n) & ~(1ull << q);
    u64 add = 0ull;
    if (W.lane == (q & 31)) {
        if (NS == 1 || q < 32) {
            add = want & ~W.t0;
            W.t0 |= want;
            W.rc0 = W.n - 1;
        } else {
            add = want & ~W.t1;
            W.t1 |= want;
            W.rc1 = W.n - 1;
        }
    }
    add = shfl_u64(add, q & 31);
    if ((add >> W.lane) & 1ull) W.cc0 += 1;
    if (NS == 2 && ((add >> (32 + W.lane)) & 1ull)) W.cc1 += 1;
    tabu_masks(W);
}

// findPivotBV (lpsol.h:552-663).  Returns the pivot ROW or -1; leaves column q in
// the scratch line.
template <int MR, int NS>
__device__ __forceinline__ int ratio_test(LP<MR, NS> &W, int q)
{
    col_to_scratch(W, q);
    const int lane = W.lane;
    const double aq = lane < MR ? W.sc[lane] : 0.0;
    const u64 tq = shfl_u64(pick<NS>(W.t0, W.t1, q >> 5), q & 31);
    const bool rowok = lane < W.m && !((tq >> W.e2b) & 1ull) && !((W.colfull >> W.e2b) & 1ull);
    const bool ok1 = rowok && !le_zero(aq); // pass 1, :571-612
    if (__any_sync(FULL, ok1)) return warp_argmin(xp_div(W.b, ok1 ? aq : 1.0), ok1);
    const bool ok2 = rowok && !is_zero(aq); // pass 2, :623-658
    if (!__any_sync(FULL, ok2)) return -1;
    return warp_argmin(xp_div(W.b, ok2 ? aq : 1.0), ok2);
}

// SIX::pivot (lpsol.h:1455-1511) on row p, entering variable q.  `have_col`: the
// scratch line already holds column q (the ratio test left it there).
template <int MR, int NS>
__device__ __forceinline__ void pivot(LP<MR, NS> &W, int p, int q, bool have_col)
{
    if (!have_col) col_to_scratch(W, q);
    const int lane = W.lane;
    const int bv = __shfl_sync(FULL, W.e2b, p);
    const double pv = W.sc[p];
    const double aq = lane < MR ? W.sc[lane] : 0.0;
    const double cq = at_col<NS>(W.c0, W.c1, q);
    const double r = xp_div(1.0, pv);
    const bool r_one = is_one(r), r_zero = is_zero(r);
    const bool cq_zero = is_zero(cq), cq_one = is_one(cq);
    double rp0, rp1;
    get_row(W, p, rp0, rp1);
    rp0 = xp_scale(rp0, r, r_one, r_zero); // :1471
    rp1 = xp_scale(rp1, r, r_one, r_zero);
    const double rhsp = xp_scale(__shfl_sync(FULL, W.b, p), r, r_one, r_zero);
    // objective row, :1496-1501
#define XPW_OBJ(s)                                            \
    {                                                         \
        double t = xp_mul(rp##s, -1.0);                       \
        t = cq_zero ? 0.0 : (cq_one ? t : xp_mul(t, cq));     \
        W.c##s = xp_add(t, W.c##s);                           \
    }
    XPW_EACH(XPW_OBJ)
#undef XPW_OBJ
    {
        double t = xp_mul(rhsp, -1.0);
        t = -t;
        t = cq_zero ? 0.0 : (cq_one ? t : xp_mul(t, cq));
        W.cr = xp_add(t, W.cr);
    }
#pragma unroll
    for (int i = 0; i < MR; i += 2) { // rank-1 elimination, :1481-1490 (row p is overwritten below)
        const double2 f2 = *reinterpret_cast<const double2 *>(W.sc + i);
        const double f0 = -f2.x, f1 = -f2.y; // coeff_of_nv = -eq(i, nv), :1485
        W.a0[i] = xp_add(W.a0[i], xp_mul(f0, rp0));
        W.a0[i + 1] = xp_add(W.a0[i + 1], xp_mul(f1, rp0));
        if (NS == 2) {
            W.a1[i] = xp_add(W.a1[i], xp_mul(f0, rp1));
            W.a1[i + 1] = xp_add(W.a1[i + 1], xp_mul(f1, rp1));
        }
    }
    W.b = (lane == p) ? rhsp : xp_add(W.b, xp_mul(-aq, rhsp));
    set_row(W, p, rp0, rp1);
    // basis maps, :1504-1510
    W.nvm = (W.nvm & ~(1ull << q)) | (1ull << bv);
    if (lane == p) W.e2b = q;
    if (lane == q) W.b2e0 = p;
    if (lane == bv) W.b2e0 = -1;
    if (NS == 2) {
        if (32 + lane == q) W.b2e1 = p;
        if (32 + lane == bv) W.b2e1 = -1;
    }
    W.pivots++;
}

// Optimal exit: sol from the basis + is_feasible (lpsol.h:1089-1127, :783-822), vc = -I | 0.
template <int MR, int NS>
__device__ __forceinline__ int optimal_exit(LP<MR, NS> &W)
{
    const int lane = W.lane, n = W.n;
    bool bad = false;
#define XPW_SOL(s)                                                              \
    {                                                                           \
        const int j = 32 * s + lane;                                            \
        const bool basic = j < n && !((W.nvm >> j) & 1ull);                     \
        const double v = __shfl_sync(FULL, W.b, basic ? W.b2e##s : 0);          \
        W.sol##s = basic ? v : 0.0;                                             \
        if (xp_mul(-1.0, W.sol##s) > 0.0) bad = true; /* vc(i,i)*sol(i) > vc(i,rhs) */ \
    }
    XPW_EACH(XPW_SOL)
#undef XPW_SOL
    // row sums, left to right over the basic columns (non-basic terms are exact +-0)
    double sum = 0.0;
    u64 basics = ~W.nvm & low_mask(n);
    while (basics) {
        const int j = __ffsll((long long)basics) - 1;
        basics &= basics - 1ull;
        col_to_scratch(W, j);
        const double solj = at_col<NS>(W.sol0, W.sol1, j);
        const double aij = lane < MR ? W.sc[lane] : 0.0;
        sum = xp_add(sum, xp_mul(aij, solj));
    }
    if (lane < W.m && !xp_feq(sum, W.b)) bad = true;
    return __any_sync(FULL, bad) ? XP_SIX_OPTIMAL_IS_INFEASIBLE : XP_SIX_SUCC;
}

// solveSlackForm (lpsol.h:1007-1191).  Returns the SIX status; *iters = cnt.
template <int MR, int NS>
__device__ __forceinline__ int solve_loop(LP<MR, NS> &W, uint32_t max_iter, uint32_t *iters)
{
    const int lane = W.lane;
    tabu_reset(W);
    W.sol0 = W.sol1 = 0.0; // sol.reinit, :1028
    const int n = W.n;
    uint32_t cnt = 0;
    while (cnt < max_iter) {
        int q = -1, p = -1;
        for (;;) {
            // pricing, :1054-1069
            const u64 posm = colmask<NS>(W.c0 > 0.0, W.c1 > 0.0) & W.nvm & low_mask(n);
            const u64 cand = posm & ~W.rowfull;
            const int best = cand ? __ffsll((long long)cand) - 1 : XPB_BIG;
            const int zlim = best == XPB_BIG ? n : best;
            // basic coefficients passed by the scan are forced to 0, :1059
            if (lane < zlim && !((W.nvm >> lane) & 1ull)) W.c0 = 0.0;
            if (NS == 2 && 32 + lane < zlim && !((W.nvm >> (32 + lane)) & 1ull)) W.c1 = 0.0;
            if (best == XPB_BIG) {
                if (!posm) { // optimal exit, :1089-1127
                    *iters = cnt;
                    return optimal_exit(W);
                }
                // findPivotNVandBVPair, :670-773
                bool found = false;
                for (int pass = 0; pass < 2 && !found; pass++) {
                    const bool z0 = !(W.c0 > 0.0) && is_zero(W.c0);
                    const bool z1 = !(W.c1 > 0.0) && is_zero(W.c1);
                    u64 cm = pass == 0 ? colmask<NS>(W.c0 > 0.0, W.c1 > 0.0) : colmask<NS>(z0, z1);
                    cm &= W.nvm & low_mask(n) & ~W.rowfull;
                    while (cm) {
                        const int j = __ffsll((long long)cm) - 1;
                        cm &= cm - 1ull;
                        const int r = ratio_test(W, j);
                        if (r >= 0) {
                            q = j;
                            p = r;
                            found = true;
                            break;
                        }
                    }
                }
                if (!found) {
                    *iters = cnt;
                    return XP_SIX_UNBOUND; // :1138-1141
                }
                break;
            }
            q = best;
            p = ratio_test(W, q);
            if (p >= 0) break;
            tabu_disable_nv(W, q); // :1146-1151
        }
        const int bv = __shfl_sync(FULL, W.e2b, p);
        tabu_gen_pair(W, q, bv); // :1156
        pivot(W, p, q, true);    // :1170
        cnt++;
    }
    *iters = cnt;
    return XP_SIX_TIME_OUT;
}

// lpsol.h:944-953 with FloatMat::substit (xmat.cpp:1491-1520), is_eq=false.
template <int MR, int NS>
__device__ __forceinline__ void restore_objective(LP<MR, NS> &W, const double *tg, int n_orig)
{
    const int lane = W.lane, rhs = W.n;
    W.c0 = lane < n_orig ? tg[lane] : 0.0;
    W.c1 = (NS == 2 && 32 + lane < n_orig) ? tg[32 + lane] : 0.0;
    W.cr = tg[n_orig];
    for (int i = 0; i < rhs; i++) {
        const double ci = at_col<NS>(W.c0, W.c1, i);
        if (is_zero(ci) || ((W.nvm >> i) & 1ull)) continue;
        const int row = at_col_i<NS>(W.b2e0, W.b2e1, i);
        double ex0, ex1;
        get_row(W, row, ex0, ex1);
        const double exr = __shfl_sync(FULL, W.b, row);
        const double ev = at_col<NS>(ex0, ex1, i);
        const bool skip = is_zero(ev);
        double sv = -1.0;
        if (!xp_feq(ci, ev)) sv = xp_div(-ci, ev);
        const bool s_zero = is_zero(sv), s_one = is_one(sv);
        if (!skip) {
#define XPW_SUB(s)                                                                     \
    {                                                                                  \
        const double x = s_zero ? 0.0 : (s_one ? ex##s : xp_mul(ex##s, sv));           \
        W.c##s = xp_add(x, W.c##s);                                                    \
    }
            XPW_EACH(XPW_SUB)
#undef XPW_SUB
        }
        {
            double tj = xp_mul(W.cr, -1.0);
            if (!skip) {
                const double x = s_zero ? 0.0 : (s_one ? exr : xp_mul(exr, sv));
                tj = xp_add(x, tj);
            }
            W.cr = xp_mul(tj, -1.0);
        }
    }
}

// Drop column xa: columns (xa, n) move one to the left (lpsol.h:956-986).
template <int NS, class T>
__device__ __forceinline__ void shift_left(T &x0, T &x1, int xa, int lane, T fill)
{
    const T d0 = __shfl_down_sync(FULL, x0, 1);
    if (NS == 2) {
        const T d1 = __shfl_down_sync(FULL, x1, 1);
        const T w = __shfl_sync(FULL, x1, 0);
        if (lane >= xa) x0 = lane < 31 ? d0 : w;
        if (32 + lane >= xa) x1 = lane < 31 ? d1 : fill;
    } else {
        if (lane >= xa) x0 = lane < 31 ? d0 : fill;
    }
}

template <int MR, int NS>
__device__ __forceinline__ int two_stage(LP<MR, NS> &W, const XpBatchArgs &A, const double *leq,
                                         const double *tg, int m, int n, uint32_t *iters)
{
    const int lane = W.lane;
    // stage1 decision, :1794-1803
    bool pos = false, bneg = false;
    for (int j = lane; j < n; j += 32) pos |= tg[j] > 0.0;
    const double b_in = lane < m ? leq[(size_t)lane * (n + 1) + n] : 0.0;
    bneg = lane < m && b_in < 0.0;
    pos = __any_sync(FULL, pos);
    bneg = __any_sync(FULL, bneg);
    const bool aux = !pos || bneg;
    const int xa = n;
    const int s0 = aux ? n + 1 : n;
    W.m = m;
    W.n = s0 + m;
    W.pivots = 0;
    // slack form [A | (-1) | I | b], :860-875 / :1405-1433, and the identity basis
#pragma unroll
    for (int i = 0; i < MR; i++) {
#define XPW_LOAD(s)                                                    \
    {                                                                  \
        const int j = 32 * s + lane;                                   \
        double v = 0.0;                                                \
        if (i < m) {                                                   \
            if (j < n) v = leq[(size_t)i * (n + 1) + j];               \
            else if (aux && j == xa) v = -1.0;                         \
            else if (j < W.n) v = (j - s0 == i) ? 1.0 : 0.0;           \
        }                                                              \
        W.a##s[i] = v;                                                 \
    }
        XPW_EACH(XPW_LOAD)
#undef XPW_LOAD
    }
    W.b = b_in;
    W.e2b = lane < m ? s0 + lane : 0;
    W.c1 = 0.0;
    W.b2e1 = -1;
#define XPW_INIT(s)                                                    \
    {                                                                  \
        const int j = 32 * s + lane;                                   \
        double v = 0.0;                                                \
        if (aux) {                                                     \
            if (j == xa) v = -1.0;                                     \
        } else if (j < n) v = tg[j];                                   \
        W.c##s = v;                                                    \
        W.b2e##s = (j < W.n && j >= s0) ? j - s0 : -1;                 \
    }
    XPW_EACH(XPW_INIT)
#undef XPW_INIT
    W.sol0 = W.sol1 = 0.0;
    W.cr = aux ? 0.0 : tg[n];
    W.nvm = low_mask(s0);

    if (aux) {
        // forced first pivot on the row of the first minimum constant term, :892-908
        const int prow = warp_argmin(W.b, lane < m);
        pivot(W, prow, xa, false);
        uint32_t it1 = 0;
        const int st = solve_loop(W, A.max_iter, &it1);
        if (st != XP_SIX_SUCC) return XP_SIX_NO_PRI_FEASIBLE_SOL; // :912-915
        if (!is_zero(W.cr)) return XP_SIX_NO_PRI_FEASIBLE_SOL; // :919-922
        if (!((W.nvm >> xa) & 1ull)) { // xa still basic: pivot it out, :924-941
            const int eqnum = at_col_i<NS>(W.b2e0, W.b2e1, xa);
            double ex0, ex1;
            get_row(W, eqnum, ex0, ex1);
            u64 nz = colmask<NS>(!is_zero(ex0), !is_zero(ex1));
            nz &= W.nvm & low_mask(W.n);
            if (!nz) return XP_ERR_REFERENCE_UB; // reference ASSERTs (:937)
            pivot(W, eqnum, __ffsll((long long)nz) - 1, false);
        }
        // restore the original objective by substitution (:944-953), then drop column xa and
        // re-index the maps (:956-986)
        restore_objective(W, tg, n);
        double dz = 0.0;
#pragma unroll
        for (int i = 0; i < MR; i++) {
            if (NS == 2) shift_left<NS, double>(W.a0[i], W.a1[i], xa, lane, 0.0);
            else shift_left<NS, double>(W.a0[i], dz, xa, lane, 0.0);
        }
        shift_left<NS, double>(W.c0, W.c1, xa, lane, 0.0);
        shift_left<NS, int>(W.b2e0, W.b2e1, xa, lane, -1);
        W.nvm = (W.nvm & low_mask(xa)) | ((W.nvm >> 1) & ~low_mask(xa));
        if (W.e2b > xa) W.e2b -= 1;
        W.n -= 1;
    }
    return solve_loop(W, A.max_iter, iters);
}

template <int MR, int NS>
__device__ __forceinline__ void write_out(const LP<MR, NS> &W, const XpBatchArgs &A, int k, int st,
                                          uint32_t iters)
{
    const int lane = W.lane, n = W.n;
    if (lane == 0) {
        if (A.maxv) ((double *)A.maxv)[k] = st == XP_SIX_SUCC ? W.cr : 0.0;
        if (A.status) A.status[k] = st;
        if (A.iters) A.iters[k] = iters;
        if (A.pivots) A.pivots[k] = W.pivots;
    }
    if (A.slack_sol) {
        double *o = (double *)A.slack_sol + (size_t)k * A.ldo;
        if (lane < A.ldo) o[lane] = lane < n ? W.sol0 : 0.0;
        if (NS == 2 && 32 + lane < A.ldo) o[32 + lane] = 32 + lane < n ? W.sol1 : 0.0;
        for (int j = 32 * NS + lane; j < A.ldo; j += 32) o[j] = 0.0;
    }
    if (A.tgtf_out) {
        double *o = (double *)A.tgtf_out + (size_t)k * A.ldo;
        if (lane < A.ldo) o[lane] = lane < n ? W.c0 : (lane == n ? W.cr : 0.0);
        if (NS == 2 && 32 + lane < A.ldo) o[32 + lane] = 32 + lane < n ? W.c1 : (32 + lane == n ? W.cr : 0.0);
        for (int j = 32 * NS + lane; j < A.ldo; j += 32) o[j] = j == n ? W.cr : 0.0;
    }
    if (A.eq2bv && lane < W.m) A.eq2bv[(size_t)k * A.ldm + lane] = W.e2b;
}

// Persistent warps pull LP indices from the batch's atomic queue.
template <int MR, int NS>
__global__ void __launch_bounds__(32 * WARPS) k_warp_f64(XpBatchArgs A)
{
    __shared__ __align__(16) double scratch[WARPS][32];
    LP<MR, NS> W;
    W.c1 = W.sol1 = 0.0;
    W.t1 = 0ull;
    W.rc1 = W.cc1 = 0;
    W.b2e1 = -1;
    W.lane = threadIdx.x & 31;
    W.sc = scratch[threadIdx.x >> 5];
    for (;;) {
        int k = 0;
        if (W.lane == 0) k = (int)atomicAdd(A.queue, 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= A.batch) break;
        const int m = A.ms ? A.ms[k] : A.m;
        const int n = A.ns ? A.ns[k] : A.n;
        const double *leq = (const double *)A.leq + (A.leq_off ? A.leq_off[k] : (int64_t)k * m * (n + 1));
        const double *tg = (const double *)A.tgtf + (A.tgtf_off ? A.tgtf_off[k] : (int64_t)k * (n + 1));
        uint32_t iters = 0;
        const int st = two_stage(W, A, leq, tg, m, n, &iters);
        write_out(W, A, k, st, iters);
    }
}

} // namespace xpw

#endif // __CUDACC__
