// One-WARP-per-LP exact simplex with the integer tableau in REGISTERS: the
// fraction-free twin of SIX<RMat,Rational>::TwoStageMethod (see xp_batch_i64.cu for the
// arithmetic: a_ij = N_ij / D, one common denominator per LP, 128-bit products, exact
// division through the modular inverse of D's odd part) on the layout of
// xp_batch_warp_f64.cuh: lane l owns columns l and 32+l of at most 32 rows, the constant
// column and eq2bv are kept one row per lane, the entering column is transposed through
// a per-warp scratch line, the pivot row is picked out of the register file by a
// warp-uniform branch tree, pricing / tabu table / basis maps are ballots and 64-bit
// masks.  Same decisions and the same integers as the one-CTA-per-LP kernel; tests A/B
// the two on every instantiation.  Used for dependence-feasibility systems
// (Lineq::has_solution, ~7 x 11), where a lane's share of the tableau leaves room for
// four warps per scheduler; larger shapes (config 4, 24 x 48) run at 255 registers here
// and stay on the CTA kernel (see launch_i64).  The bound is the integer pipe (two
// 64x64->128 products + a 128-bit subtract + shift + one 64-bit multiply per entry).
//
// Reference: lpsol.h as listed in xp_batch_warp_f64.cuh; Rational value semantics
// rational.cpp:229-397.
#pragma once

#include "xp_batch_warp_f64.cuh"
#include "xp_exact_arith.cuh"

#ifdef __CUDACC__

namespace xpwi {

using xpw::at_col_i;
using xpw::colmask;
using xpw::FULL;
using xpw::low_mask;
using xpw::pick;
using xpw::shfl_u64;
using xpw::u64;
typedef long long i64;
typedef __int128 i128;
constexpr int WARPS = xpw::WARPS;

template <int MR, int NS>
struct LP {
    i64 a0[MR], a1[MR]; // N(i, lane), N(i, 32 + lane)
    i64 c0, c1;         // objective numerators, own columns
    i64 sol0, sol1;     // slack solution numerators (over D), own columns
    i64 cr;             // objective constant numerator (uniform)
    i64 b;              // constant column, row `lane`
    xpx::Div dv;        // common denominator D and the constants of the division by it (uniform)
    int e2b;
    int b2e0, b2e1;
    u64 t0, t1;
    int rc0, rc1, cc0, cc1;
    u64 nvm;
    u64 rowfull, colfull;
    int m, n;
    unsigned pivots;
    int lane;
    i64 *sc; // this warp's scratch line (32 x int64, shared memory)
    static constexpr int kNS = NS;
};

__device__ __forceinline__ i64 shfl_i64(i64 x, int src) { return (i64)shfl_u64((u64)x, src); }
template <int NS>
__device__ __forceinline__ i64 at_col(i64 x0, i64 x1, int j)
{
    return shfl_i64(pick<NS>(x0, x1, j >> 5), j & 31);
}

using xpx::ff_div;
using xpx::gcd64;

template <int MR, int NS>
__device__ __forceinline__ void col_to_scratch(LP<MR, NS> &W, int q)
{
    __syncwarp();
    if (W.lane == (q & 31)) {
        if (NS == 1 || q < 32) {
#pragma unroll
            for (int i = 0; i < MR; i += 2)
                *reinterpret_cast<longlong2 *>(W.sc + i) = make_longlong2(W.a0[i], W.a0[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < MR; i += 2)
                *reinterpret_cast<longlong2 *>(W.sc + i) = make_longlong2(W.a1[i], W.a1[i + 1]);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ i64 opaque(i64 x)
{
    i64 y;
    asm volatile("mov.b64 %0, %1;" : "=l"(y) : "l"(x));
    return y;
}
template <int MR, int NS>
__device__ __forceinline__ void get_row(const LP<MR, NS> &W, int p, i64 &r0, i64 &r1)
{
    r0 = 0;
    r1 = 0;
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
__device__ __forceinline__ void set_row(LP<MR, NS> &W, int p, i64 r0, i64 r1)
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

// Warp arg-min of the fractions num/den (den > 0) over the lanes with `on`, first strict
// minimum = lowest row on ties (lpsol.h:603-611); comparisons are 128-bit cross products,
// the value semantics of Rational::operator< (rational.cpp:229-270).
__device__ __forceinline__ int warp_argmin_frac(i64 num, i64 den, bool on, int lane)
{
    int idx = on ? lane : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const i64 n2 = (i64)__shfl_xor_sync(FULL, (long long)num, o);
        const i64 d2 = (i64)__shfl_xor_sync(FULL, (long long)den, o);
        const int i2 = __shfl_xor_sync(FULL, idx, o);
        bool take = false;
        if (i2 >= 0) {
            if (idx < 0) take = true;
            else {
                const i128 l = (i128)n2 * den, r = (i128)num * d2; // other < mine ?
                take = l < r || (l == r && i2 < idx);
            }
        }
        if (take) {
            num = n2;
            den = d2;
            idx = i2;
        }
    }
    return idx;
}

// findPivotBV (lpsol.h:552-663).  Returns the pivot ROW or -1; leaves column q in the scratch line.
template <int MR, int NS>
__device__ __forceinline__ int ratio_test(LP<MR, NS> &W, int q)
{
    col_to_scratch(W, q);
    const int lane = W.lane;
    const i64 aq = lane < MR ? W.sc[lane] : 0;
    const u64 tq = shfl_u64(pick<NS>(W.t0, W.t1, q >> 5), q & 31);
    const bool rowok = lane < W.m && !((tq >> W.e2b) & 1ull) && !((W.colfull >> W.e2b) & 1ull);
    // v = rhs / coeff as a sign-normalised fraction (the common D cancels)
    const i64 num = aq < 0 ? -W.b : W.b, den = aq < 0 ? -aq : aq;
    const bool ok1 = rowok && aq > 0; // pass 1, :571-612
    if (__any_sync(FULL, ok1)) return warp_argmin_frac(num, den, ok1, lane);
    const bool ok2 = rowok && aq != 0; // pass 2, :623-658
    if (!__any_sync(FULL, ok2)) return -1;
    return warp_argmin_frac(num, den, ok2, lane);
}

// SIX::pivot (lpsol.h:1455-1511), integer preserving.  Returns 0 or XP_ERR_OVERFLOW.
template <int MR, int NS>
__device__ __forceinline__ int pivot(LP<MR, NS> &W, int p, int q, bool have_col)
{
    if (!have_col) col_to_scratch(W, q);
    const int lane = W.lane;
    const int bv = __shfl_sync(FULL, W.e2b, p);
    const i64 P = W.sc[p];
    const i64 aql = lane < MR ? W.sc[lane] : 0;
    const i64 cq = at_col<NS>(W.c0, W.c1, q);
    const xpx::Div dv = W.dv;
    const i64 aP = P < 0 ? -P : P;
    const bool sneg = P < 0;
    const i64 scq = sneg ? -cq : cq;
    i64 rp0, rp1;
    get_row(W, p, rp0, rp1);
    const i64 rhsp = shfl_i64(W.b, p);
    bool ovf = false;
    // objective row: Nt_j <- (Nt_j*|P| - s*Nt_q*N_pj) / D, constant with + (:1496-1501)
    W.c0 = ff_div((i128)W.c0 * aP - (i128)scq * rp0, dv, ovf);
    if (NS == 2) W.c1 = ff_div((i128)W.c1 * aP - (i128)scq * rp1, dv, ovf);
    W.cr = ff_div((i128)W.cr * aP + (i128)scq * rhsp, dv, ovf);
#pragma unroll
    for (int i = 0; i < MR; i += 2) { // N_ij <- (N_ij*|P| - s*N_iq*N_pj) / D; row p comes out as 0 and is set below
        const longlong2 f2 = *reinterpret_cast<const longlong2 *>(W.sc + i);
        const i64 f0 = sneg ? -f2.x : f2.x, f1 = sneg ? -f2.y : f2.y;
        W.a0[i] = ff_div((i128)W.a0[i] * aP - (i128)f0 * rp0, dv, ovf);
        W.a0[i + 1] = ff_div((i128)W.a0[i + 1] * aP - (i128)f1 * rp0, dv, ovf);
        if (NS == 2) {
            W.a1[i] = ff_div((i128)W.a1[i] * aP - (i128)f0 * rp1, dv, ovf);
            W.a1[i + 1] = ff_div((i128)W.a1[i + 1] * aP - (i128)f1 * rp1, dv, ovf);
        }
        asm volatile("" ::: "memory"); // keep the multiplier loads with their rows (register pressure)
    }
    {
        const i64 fl = sneg ? -aql : aql;
        const i64 nb = ff_div((i128)W.b * aP - (i128)fl * rhsp, dv, ovf);
        W.b = (lane == p) ? (sneg ? -rhsp : rhsp) : nb;
    }
    set_row(W, p, sneg ? -rp0 : rp0, sneg ? -rp1 : rp1); // row p: N_pj <- s*N_pj
    W.dv.set((u64)aP);
    W.nvm = (W.nvm & ~(1ull << q)) | (1ull << bv);
    if (lane == p) W.e2b = q;
    if (lane == q) W.b2e0 = p;
    if (lane == bv) W.b2e0 = -1;
    if (NS == 2) {
        if (32 + lane == q) W.b2e1 = p;
        if (32 + lane == bv) W.b2e1 = -1;
    }
    W.pivots++;
    return __any_sync(FULL, ovf) ? XP_ERR_OVERFLOW : 0;
}

// Optimal exit: in exact arithmetic the row-sum half of is_feasible (lpsol.h:805-814) is an
// identity, so the sign test on the basic values decides (:798-802).
template <int MR, int NS>
__device__ __forceinline__ int optimal_exit(LP<MR, NS> &W)
{
    const int lane = W.lane, n = W.n;
    bool bad = false;
#define XPW_SOL(s)                                                       \
    {                                                                    \
        const int j = 32 * s + lane;                                     \
        const bool basic = j < n && !((W.nvm >> j) & 1ull);              \
        const i64 v = shfl_i64(W.b, basic ? W.b2e##s : 0);               \
        W.sol##s = basic ? v : 0;                                        \
        if (W.sol##s < 0) bad = true;                                    \
    }
    XPW_EACH(XPW_SOL)
#undef XPW_SOL
    return __any_sync(FULL, bad) ? XP_SIX_OPTIMAL_IS_INFEASIBLE : XP_SIX_SUCC;
}

// solveSlackForm (lpsol.h:1007-1191).  Returns the SIX status or XP_ERR_OVERFLOW; *iters = cnt.
template <int MR, int NS>
__device__ __forceinline__ int solve_loop(LP<MR, NS> &W, uint32_t max_iter, uint32_t *iters)
{
    const int lane = W.lane;
    xpw::tabu_reset(W);
    W.sol0 = W.sol1 = 0;
    const int n = W.n;
    uint32_t cnt = 0;
    while (cnt < max_iter) {
        int q = -1, p = -1;
        for (;;) {
            const u64 posm = colmask<NS>(W.c0 > 0, W.c1 > 0) & W.nvm & low_mask(n); // D > 0
            const u64 cand = posm & ~W.rowfull;
            const int best = cand ? __ffsll((long long)cand) - 1 : XPB_BIG;
            const int zlim = best == XPB_BIG ? n : best;
            if (lane < zlim && !((W.nvm >> lane) & 1ull)) W.c0 = 0; // :1059
            if (NS == 2 && 32 + lane < zlim && !((W.nvm >> (32 + lane)) & 1ull)) W.c1 = 0;
            if (best == XPB_BIG) {
                if (!posm) {
                    *iters = cnt;
                    return optimal_exit(W);
                }
                bool found = false; // findPivotNVandBVPair, :670-773
                for (int pass = 0; pass < 2 && !found; pass++) {
                    u64 cm = pass == 0 ? colmask<NS>(W.c0 > 0, W.c1 > 0) : colmask<NS>(W.c0 == 0, W.c1 == 0);
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
                    return XP_SIX_UNBOUND;
                }
                break;
            }
            q = best;
            p = ratio_test(W, q);
            if (p >= 0) break;
            xpw::tabu_disable_nv(W, q);
        }
        const int bv = __shfl_sync(FULL, W.e2b, p);
        xpw::tabu_gen_pair(W, q, bv);
        const int rc = pivot(W, p, q, true);
        if (rc) {
            *iters = cnt;
            return rc;
        }
        cnt++;
    }
    *iters = cnt;
    return XP_SIX_TIME_OUT;
}

// lpsol.h:944-953 in exact arithmetic: each basic variable i with c_i != 0 contributes
// -c_i * row(i) (+ on the constant column); basic columns are unit vectors, so the
// substitutions commute and c_i is the input value.
template <int MR, int NS>
__device__ __forceinline__ int restore_objective(LP<MR, NS> &W, const i64 *tg, int n_orig)
{
    const int lane = W.lane;
    const i64 D = (i64)W.dv.D;
    i128 acc0 = lane < n_orig ? (i128)tg[lane] * D : (i128)0;
    i128 acc1 = (NS == 2 && 32 + lane < n_orig) ? (i128)tg[32 + lane] * D : (i128)0;
    i128 accr = (i128)tg[n_orig] * D;
    for (int i = 0; i < n_orig; i++) {
        const i64 ci = tg[i];
        if (ci == 0 || ((W.nvm >> i) & 1ull)) continue;
        const int row = at_col_i<NS>(W.b2e0, W.b2e1, i);
        i64 ex0, ex1;
        get_row(W, row, ex0, ex1);
        const i64 exr = shfl_i64(W.b, row);
        acc0 -= (i128)ci * ex0;
        acc1 -= (i128)ci * ex1;
        accr += (i128)ci * exr;
    }
    const i128 lim = (i128)0x7fffffffffffffffLL;
    bool ovf = acc0 > lim || acc0 < -lim || acc1 > lim || acc1 < -lim || accr > lim || accr < -lim;
    W.c0 = (i64)acc0;
    W.c1 = (i64)acc1;
    W.cr = (i64)accr;
    return __any_sync(FULL, ovf) ? XP_ERR_OVERFLOW : 0;
}

template <int NS>
__device__ __forceinline__ void shift_left_i64(i64 &x0, i64 &x1, int xa, int lane)
{
    const i64 d0 = (i64)__shfl_down_sync(FULL, (long long)x0, 1);
    if (NS == 2) {
        const i64 d1 = (i64)__shfl_down_sync(FULL, (long long)x1, 1);
        const i64 w = (i64)__shfl_sync(FULL, (long long)x1, 0);
        if (lane >= xa) x0 = lane < 31 ? d0 : w;
        if (32 + lane >= xa) x1 = lane < 31 ? d1 : 0;
    } else {
        if (lane >= xa) x0 = lane < 31 ? d0 : 0;
    }
}

template <int MR, int NS>
__device__ __forceinline__ int two_stage(LP<MR, NS> &W, const XpBatchArgs &A, const i64 *leq, const i64 *tg,
                                         int m, int n, uint32_t *iters)
{
    const int lane = W.lane;
    bool pos = false, bneg = false; // stage1 decision, :1794-1803
    for (int j = lane; j < n; j += 32) pos |= tg[j] > 0;
    const i64 b_in = lane < m ? leq[(size_t)lane * (n + 1) + n] : 0;
    bneg = lane < m && b_in < 0;
    pos = __any_sync(FULL, pos);
    bneg = __any_sync(FULL, bneg);
    const bool aux = !pos || bneg;
    const int xa = n;
    const int s0 = aux ? n + 1 : n;
    W.m = m;
    W.n = s0 + m;
    W.pivots = 0;
    W.dv.set(1ull);
#pragma unroll
    for (int i = 0; i < MR; i++) {
#define XPW_LOAD(s)                                                    \
    {                                                                  \
        const int j = 32 * s + lane;                                   \
        i64 v = 0;                                                     \
        if (i < m) {                                                   \
            if (j < n) v = leq[(size_t)i * (n + 1) + j];               \
            else if (aux && j == xa) v = -1;                           \
            else if (j < W.n) v = (j - s0 == i) ? 1 : 0;               \
        }                                                              \
        W.a##s[i] = v;                                                 \
    }
        XPW_EACH(XPW_LOAD)
#undef XPW_LOAD
    }
    W.b = b_in;
    W.e2b = lane < m ? s0 + lane : 0;
    W.c1 = 0;
    W.b2e1 = -1;
#define XPW_INIT(s)                                                    \
    {                                                                  \
        const int j = 32 * s + lane;                                   \
        i64 v = 0;                                                     \
        if (aux) {                                                     \
            if (j == xa) v = -1;                                       \
        } else if (j < n) v = tg[j];                                   \
        W.c##s = v;                                                    \
        W.b2e##s = (j < W.n && j >= s0) ? j - s0 : -1;                 \
    }
    XPW_EACH(XPW_INIT)
#undef XPW_INIT
    W.sol0 = W.sol1 = 0;
    W.cr = aux ? 0 : tg[n];
    W.nvm = low_mask(s0);

    if (aux) {
        // forced first pivot on the row of the first minimum constant term, :892-908
        const int prow = warp_argmin_frac(W.b, 1, lane < m, lane);
        int rc = pivot(W, prow, xa, false);
        if (rc) return rc;
        uint32_t it1 = 0;
        const int st = solve_loop(W, A.max_iter, &it1);
        if (st < 0) return st;
        if (st != XP_SIX_SUCC) return XP_SIX_NO_PRI_FEASIBLE_SOL; // :912-915
        if (W.cr != 0) return XP_SIX_NO_PRI_FEASIBLE_SOL;          // :919-922
        if (!((W.nvm >> xa) & 1ull)) {                             // xa still basic, :924-941
            const int eqnum = at_col_i<NS>(W.b2e0, W.b2e1, xa);
            i64 ex0, ex1;
            get_row(W, eqnum, ex0, ex1);
            u64 nz = colmask<NS>(ex0 != 0, ex1 != 0);
            nz &= W.nvm & low_mask(W.n);
            if (!nz) return XP_ERR_REFERENCE_UB;
            rc = pivot(W, eqnum, __ffsll((long long)nz) - 1, false);
            if (rc) return rc;
        }
        rc = restore_objective(W, tg, n); // :944-953
        if (rc) return rc;
        i64 dz = 0;
#pragma unroll
        for (int i = 0; i < MR; i++) { // drop column xa, :956-986
            if (NS == 2) shift_left_i64<NS>(W.a0[i], W.a1[i], xa, lane);
            else shift_left_i64<NS>(W.a0[i], dz, xa, lane);
        }
        shift_left_i64<NS>(W.c0, W.c1, xa, lane);
        xpw::shift_left<NS, int>(W.b2e0, W.b2e1, xa, lane, -1);
        W.nvm = (W.nvm & low_mask(xa)) | ((W.nvm >> 1) & ~low_mask(xa));
        if (W.e2b > xa) W.e2b -= 1;
        W.n -= 1;
    }
    return solve_loop(W, A.max_iter, iters);
}

__device__ __forceinline__ void put_frac(i64 num, i64 D, i64 *onum, i64 *oden)
{ // reduced num/den with a positive denominator; 0 is 0/1
    i64 den = 1;
    if (num != 0) {
        const i64 g = gcd64(num, D);
        num /= g;
        den = D / g;
    }
    *onum = num;
    if (oden) *oden = den;
}

template <int MR, int NS>
__device__ __forceinline__ void write_out(const LP<MR, NS> &W, const XpBatchArgs &A, int k, int st, uint32_t iters)
{
    const int lane = W.lane, n = W.n;
    const i64 D = (i64)W.dv.D;
    if (lane == 0) {
        if (A.maxv) put_frac(st == XP_SIX_SUCC ? W.cr : 0, D, (i64 *)A.maxv + 2 * (size_t)k, (i64 *)A.maxv + 2 * (size_t)k + 1);
        if (A.status) A.status[k] = st;
        if (A.iters) A.iters[k] = iters;
        if (A.pivots) A.pivots[k] = W.pivots;
    }
    for (int j = lane; j < A.ldo; j += 32) {
        const size_t o = (size_t)k * A.ldo + j;
        i64 sv = 0, cv = 0;
        if (j < 32) {
            sv = j < n ? W.sol0 : 0;
            cv = j < n ? W.c0 : (j == n ? W.cr : 0);
        } else if (NS == 2 && j < 64) {
            sv = j < n ? W.sol1 : 0;
            cv = j < n ? W.c1 : (j == n ? W.cr : 0);
        } else {
            cv = j == n ? W.cr : 0;
        }
        if (A.slack_sol) put_frac(sv, D, (i64 *)A.slack_sol + o, A.slack_sol2 ? (i64 *)A.slack_sol2 + o : nullptr);
        if (A.tgtf_out) put_frac(cv, D, (i64 *)A.tgtf_out + o, A.tgtf_out2 ? (i64 *)A.tgtf_out2 + o : nullptr);
    }
    if (A.eq2bv && lane < W.m) A.eq2bv[(size_t)k * A.ldm + lane] = W.e2b;
}

template <int MR, int NS>
__global__ void __launch_bounds__(32 * WARPS) k_warp_i64(XpBatchArgs A)
{
    __shared__ __align__(16) i64 scratch[WARPS][32];
    LP<MR, NS> W;
    W.c1 = W.sol1 = 0;
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
        const i64 *leq = (const i64 *)A.leq + (A.leq_off ? A.leq_off[k] : (int64_t)k * m * (n + 1));
        const i64 *tg = (const i64 *)A.tgtf + (A.tgtf_off ? A.tgtf_off[k] : (int64_t)k * (n + 1));
        uint32_t iters = 0;
        const int st = two_stage(W, A, leq, tg, m, n, &iters);
        write_out(W, A, k, st, iters);
    }
}

} // namespace xpwi

#endif // __CUDACC__
