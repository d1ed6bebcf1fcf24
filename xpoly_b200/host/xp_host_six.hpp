// Host-side (C++) half of the drop-in: everything SIX<Mat,T>::maxm / minm do
// AROUND the hot loop -- verify, normalize (equalities -> inequalities, free
// variable split), explicit dual for minm, calcFinalSolution -- plus the
// depth-first branch & bound of MIP<Mat,T>.  The hot loop itself
// (TwoStageMethod: slack form, phase 1, solveSlackForm) runs on the GPU and is
// reached through the `solver` callback, so this header has no CUDA in it.
//
// Reference: /root/reference/src/com/lpsol.h -- verify :1516-1558,
// convertEq2Ineq :1196-1278, normalize :1289-1394, calcDualMaxm :1585-1655,
// minm :1661-1732, calcFinalSolution :1850-1899, maxm :1992-2033,
// MIP::RecusivePart :2426-2612, MIP::is_satisfying :2363-2408.
//
// Two scalar policies: F64 (the reference's Float: IEEE double, tolerant ==
// with eps 1e-17, flty.cpp:41-131) and Q (exact rationals over int64 with
// 128-bit intermediates: the value semantics of the reference's Rational
// wherever that one stays exact, rational.cpp:229-397).
#pragma once

#include <stdint.h>

#include <cmath>
#include <cstdlib>
#include <functional>
#include <vector>

#include "../../include/xpoly_b200.h"

namespace xph {

constexpr double kEps = 0.00000000000000001; // INFINITESIMAL, flty.h:46

// ------------------------------------------------------------------ F64 policy
struct F64 {
    typedef double T;
    static bool &overflow() // FP64 never overflows into an error; kept for a uniform interface
    {
        static thread_local bool f = false;
        return f;
    }
    static T zero() { return 0.0; }
    static T from_int(long long i) { return (double)i; }
    static T add(T a, T b) { return a + b; }
    static T sub(T a, T b) { return a - b; }
    static T mul(T a, T b) { return a * b; }
    static T div(T a, T b) { return a / b; }
    static T neg(T a) { return -a; }
    static bool gt(T a, T b) { return a > b; }
    static bool lt(T a, T b) { return a < b; }
    static bool eq(T a, T b)
    { // Float::operator==, flty.cpp:41-58
        if ((a > 0 && b < 0) || (a < 0 && b > 0)) return false;
        if (a < 0) a = -a;
        if (b < 0) b = -b;
        if ((a == 0.0 && b <= kEps) || (b == 0.0 && a <= kEps)) return true;
        if (a > b) return (a - b) <= kEps;
        return (b - a) <= kEps;
    }
    static bool le(T a, T b) { return a < b || eq(a, b); }
    static bool ge(T a, T b) { return a > b || eq(a, b); }
    static T reduce(T a) { return a; }
    static bool is_int(T f)
    { // Float::is_int, flty.cpp:182-201
        double av = f < 0 ? -f : f;
        long long iv = (long long)av;
        if ((av - (double)iv) < kEps) return true;
        if (((double)(iv + 1) - av) < kEps) return true;
        return false;
    }
    static int trunc(T a) { return (int)a; } // Float::typecast2int, flty.h:85-88
    static bool ok() { return true; }
    static bool div_by_zero_is_ub(T a) { return a == 0.0; } // 1/0 = inf, then 0*inf = NaN downstream
};

// -------------------------------------------------------------------- Q policy
struct QVal {
    long long num, den; // reduced, den > 0
};
struct Q {
    typedef QVal T;
    typedef __int128 i128;
    static bool &overflow()
    {
        static thread_local bool f = false;
        return f;
    }
    static long long gcdll(long long a, long long b)
    {
        unsigned long long x = a < 0 ? 0ULL - (unsigned long long)a : (unsigned long long)a;
        unsigned long long y = b < 0 ? 0ULL - (unsigned long long)b : (unsigned long long)b;
        while (y) {
            unsigned long long t = x % y;
            x = y;
            y = t;
        }
        return (long long)x;
    }
    static i128 gcd128(i128 a, i128 b)
    {
        if (a < 0) a = -a;
        if (b < 0) b = -b;
        while (b) {
            i128 t = a % b;
            a = b;
            b = t;
        }
        return a;
    }
    static T make(i128 n, i128 d)
    {
        if (n == 0) return T{0, 1};
        if (d < 0) {
            n = -n;
            d = -d;
        }
        i128 g = gcd128(n, d);
        n /= g;
        d /= g;
        const i128 lim = (i128)0x7fffffffffffffffLL;
        if (n > lim || n < -lim || d > lim) {
            overflow() = true;
            return T{0, 1};
        }
        return T{(long long)n, (long long)d};
    }
    static T zero() { return T{0, 1}; }
    static T from_int(long long i) { return T{i, 1}; }
    static T add(T a, T b) { return make((i128)a.num * b.den + (i128)b.num * a.den, (i128)a.den * b.den); }
    static T neg(T a) { return T{-a.num, a.den}; }
    static T sub(T a, T b) { return add(a, neg(b)); }
    static T mul(T a, T b) { return make((i128)a.num * b.num, (i128)a.den * b.den); }
    static T div(T a, T b) { return make((i128)a.num * b.den, (i128)a.den * b.num); }
    static bool gt(T a, T b) { return (i128)a.num * b.den > (i128)b.num * a.den; }
    static bool lt(T a, T b) { return (i128)a.num * b.den < (i128)b.num * a.den; }
    static bool eq(T a, T b) { return a.num == b.num && a.den == b.den; }
    static bool le(T a, T b) { return !gt(a, b); }
    static bool ge(T a, T b) { return !lt(a, b); }
    static T reduce(T a) { return a; }
    static bool is_int(T a) { return a.den == 1; } // RMat::is_imat, xmat.cpp:603
    static int trunc(T a) { return (int)(a.num / a.den); } // Rational::typecast2int
    static bool ok() { return !overflow(); }
    static bool div_by_zero_is_ub(T a) { return a.num == 0; } // Rational 1/0: den = 0 garbage
};

// ------------------------------------------------------------------- matrices
template <class P>
struct Mat {
    typedef typename P::T T;
    int r = 0, c = 0;
    std::vector<T> a;
    Mat() {}
    Mat(int rows, int cols) : r(rows), c(cols), a((size_t)rows * cols, P::zero()) {}
    T &at(int i, int j) { return a[(size_t)i * c + j]; }
    const T &at(int i, int j) const { return a[(size_t)i * c + j]; }
    bool empty() const { return r == 0 || c == 0; }

    void insert_cols(int cidx, int cnum)
    { // Matrix::insertColumnsBefore, matt.h:2823-2844
        if (cnum == 0) return;
        Mat n(r, c + cnum);
        for (int i = 0; i < r; i++) {
            for (int j = 0; j < cidx; j++) n.at(i, j) = at(i, j);
            for (int j = cidx; j < c; j++) n.at(i, j + cnum) = at(i, j);
        }
        *this = n;
    }
    void del_col(int col)
    {
        Mat n(r, c - 1);
        for (int i = 0; i < r; i++) {
            for (int j = 0; j < col; j++) n.at(i, j) = at(i, j);
            for (int j = col + 1; j < c; j++) n.at(i, j - 1) = at(i, j);
        }
        *this = n;
    }
    void grow_rows(int cnt, int cols_if_empty)
    {
        if (cnt == 0) return;
        if (r == 0) c = cols_if_empty;
        a.resize((size_t)(r + cnt) * c, P::zero());
        r += cnt;
    }
    // Matrix::mulOfRow, matt.h:1352-1368 (v == 1: no-op; v == 0: zero the row)
    void mul_row(int row, T v)
    {
        if (P::eq(v, P::from_int(1))) return;
        if (P::eq(v, P::zero())) {
            for (int j = 0; j < c; j++) at(row, j) = P::zero();
            return;
        }
        for (int j = 0; j < c; j++) at(row, j) = P::mul(at(row, j), v);
    }
    void mul_all(T v)
    { // Matrix::mul, matt.h:1330-1347
        if (P::eq(v, P::zero())) {
            for (auto &x : a) x = P::zero();
            return;
        }
        if (P::eq(v, P::from_int(1))) return;
        for (auto &x : a) x = P::mul(x, v);
    }
    void mul_cols(int from, int to, T v)
    { // Matrix::mulOfColumns, matt.h:1414-1432
        if (P::eq(v, P::from_int(1))) return;
        for (int j = from; j <= to; j++)
            for (int i = 0; i < r; i++)
                at(i, j) = P::eq(v, P::zero()) ? P::zero() : P::mul(at(i, j), v);
    }
    void add_row(int to, const Mat &s, int from)
    { // Matrix::addRowToRow, matt.h:1449-1460
        for (int j = 0; j < c; j++) at(to, j) = P::add(s.at(from, j), at(to, j));
    }
    bool col_all_eq(int col, T v) const
    {
        for (int i = 0; i < r; i++)
            if (!P::eq(at(i, col), v)) return false;
        return true;
    }
    Mat row_copy(int row) const
    {
        Mat n(1, c);
        for (int j = 0; j < c; j++) n.at(0, j) = at(row, j);
        return n;
    }
};

// What TwoStageMethod hands back (lpsol.h:291-301), for a normalised LP.
template <class P>
struct TwoStageResult {
    int status = XP_ERR_CUDA;
    typename P::T maxv;               // tgtf[rhs] of the final tableau (:1119)
    std::vector<typename P::T> slack_sol; // n + m + 1 entries
    std::vector<typename P::T> tgtf;      // final objective row, n + m + 1 entries
    std::vector<int32_t> eq2bv;           // m entries
};

// solver(leq m x (n+1), tgtf 1 x (n+1), max_iter) -> TwoStageResult; x >= 0.
template <class P>
using TwoStageFn = std::function<TwoStageResult<P>(const Mat<P> &, const Mat<P> &, uint32_t)>;

// convertEq2Ineq, lpsol.h:1196-1278, including the mis-indexed tp.get(0, m) at
// :1232.  Returns 0 or XP_ERR_REFERENCE_UB.
template <class P>
int eq_to_ineq(Mat<P> &leq, const Mat<P> &eq, int rhs_idx)
{
    typedef typename P::T T;
    if (eq.empty()) return 0;
    std::vector<char> removed(eq.r, 0);
    int eq_count = eq.r;
    if (!leq.empty()) {
        for (int j = 0; j < rhs_idx; j++) {
            int nnz = 0, pos = 0;
            for (int i = 0; i < eq.r; i++) {
                if (removed[i]) continue;
                if (!P::eq(eq.at(i, j), P::zero())) {
                    nnz++;
                    pos = i;
                }
            }
            if (nnz != 1) continue;
            removed[pos] = 1;
            eq_count--;
            for (int mm = 0; mm < leq.r; mm++) {
                T v = leq.at(mm, j);
                if (P::eq(v, P::zero())) continue;
                Mat<P> tp = eq.row_copy(pos);
                if (mm >= tp.c) return XP_ERR_REFERENCE_UB; // reference reads out of bounds
                T tpv = tp.at(0, mm); // sic: the row counter is used as a column index
                if (!P::eq(tpv, P::from_int(1))) {
                    if (P::div_by_zero_is_ub(tpv)) return XP_ERR_REFERENCE_UB;
                    tp.mul_row(0, P::div(P::from_int(1), tpv));
                }
                tp.mul_row(0, v);
                leq.at(mm, j) = P::zero();
                for (int k = rhs_idx; k < tp.c; k++) tp.at(0, k) = P::neg(tp.at(0, k));
                leq.add_row(mm, tp, 0);
            }
        }
    }
    if (eq_count > 0) { // :1252-1267
        int c0 = leq.r;
        leq.grow_rows(eq_count * 2, eq.c);
        for (int i = 0; i < eq.r; i++) {
            if (removed[i]) continue;
            for (int j = 0; j < eq.c; j++) leq.at(c0, j) = eq.at(i, j);
            leq.mul_row(c0, P::from_int(-1));
            for (int j = 0; j < eq.c; j++) leq.at(c0 + 1, j) = eq.at(i, j);
            c0 += 2;
        }
    }
    return 0;
}

template <class P>
struct Normalized {
    Mat<P> leq, tgtf;
    int rhs_idx = 0;
    std::vector<int32_t> vcmap; // triples {real, dummy1, dummy2}, lpsol.h:1376-1378
    bool std_vc = true;         // every variable constraint is -x <= 0
    // newvc(i,i) and newvc(i,rhs) for the rhs_idx normalised variables: the only entries of the
    // variable constraints the solver reads (is_feasible, lpsol.h:798-802)
    std::vector<typename P::T> vc_diag, vc_rhs;
};

// normalize, lpsol.h:1289-1394.  vc: n x (n+1).
template <class P>
int normalize(Normalized<P> &N, const Mat<P> &vc, const Mat<P> &eq, const Mat<P> &leq,
              const Mat<P> &tgtf, int m_rhs_idx)
{
    const int vars = m_rhs_idx;
    Mat<P> tmpleq = leq;
    int err = eq_to_ineq(tmpleq, eq, m_rhs_idx);
    if (err) return err;
    int grow = 0;
    for (int i = 0; i < vars; i++)
        if (vc.col_all_eq(i, P::zero())) grow++;
    for (int i = 0; i < vars; i++) {
        if (vc.col_all_eq(i, P::zero())) continue;
        // the GPU feasibility test assumes vc(i,i) < 0 and vc(i,rhs) == 0
        if (!(i < vc.r) || !P::lt(vc.at(i, i), P::zero()) || !P::eq(vc.at(i, vc.c - 1), P::zero()))
            N.std_vc = false;
    }
    N.leq = tmpleq;
    N.leq.insert_cols(m_rhs_idx, grow);
    N.tgtf = tgtf;
    N.tgtf.insert_cols(m_rhs_idx, grow);
    int last = m_rhs_idx - 1;
    N.vcmap.clear();
    for (int i = 0; i < vars; i++) { // :1365-1392
        if (!vc.col_all_eq(i, P::zero())) continue;
        N.vcmap.push_back(i);
        N.vcmap.push_back(i);
        N.vcmap.push_back(last + 1);
        for (int r = 0; r < N.leq.r; r++) N.leq.at(r, last + 1) = tmpleq.at(r, i);
        N.leq.mul_cols(last + 1, last + 1, P::from_int(-1));
        N.tgtf.at(0, last + 1) = tgtf.at(0, i);
        N.tgtf.mul_cols(last + 1, last + 1, P::from_int(-1));
        last++;
    }
    N.rhs_idx = last + 1;
    // newvc as normalize leaves it (:1341-1373): the caller's square part and constant column for
    // the original variables, -1 / 0 for both halves of a split free variable and for every dummy
    N.vc_diag.assign(N.rhs_idx, P::from_int(-1));
    N.vc_rhs.assign(N.rhs_idx, P::zero());
    for (int i = 0; i < vars; i++) {
        if (vc.col_all_eq(i, P::zero())) continue;
        N.vc_diag[i] = i < vc.r ? vc.at(i, i) : P::zero();
        N.vc_rhs[i] = i < vc.r ? vc.at(i, vc.c - 1) : P::zero();
    }
    return 0;
}

// The variable-constraint half of is_feasible (lpsol.h:798-802) on a solution row: true if some
// vc(i,i) * sol(i) > vc(i,rhs).  slack / auxiliary variables beyond `diag` are -x <= 0.
template <class P>
bool violates_vc(const std::vector<typename P::T> &sol, int rhs_idx, const std::vector<typename P::T> &diag,
                 const std::vector<typename P::T> &rhs)
{
    for (int i = 0; i < rhs_idx && i < (int)sol.size(); i++) {
        const typename P::T d = i < (int)diag.size() ? diag[i] : P::from_int(-1);
        const typename P::T r = i < (int)rhs.size() ? rhs[i] : P::zero();
        if (P::lt(r, P::mul(d, sol[i]))) return true;
    }
    return false;
}

// calcFinalSolution, lpsol.h:1850-1899.
template <class P>
void final_solution(std::vector<typename P::T> &sol, typename P::T &v,
                    std::vector<typename P::T> &slack_sol, const std::vector<int32_t> &vcmap,
                    const Mat<P> &orig_tgtf, int m_rhs_idx)
{
    for (size_t i = 0; i + 2 < vcmap.size(); i += 3)
        slack_sol[vcmap[i]] = P::sub(slack_sol[vcmap[i + 1]], slack_sol[vcmap[i + 2]]);
    sol.assign(orig_tgtf.c, P::zero());
    for (int i = 0; i < m_rhs_idx; i++) sol[i] = slack_sol[i];
    for (int k = m_rhs_idx; k < orig_tgtf.c; k++) sol[k] = P::from_int(1);
    v = P::zero();
    for (int j = 0; j < orig_tgtf.c; j++) v = P::add(v, P::mul(sol[j], orig_tgtf.at(0, j)));
}

template <class P>
Mat<P> default_vc(int n)
{
    Mat<P> V(n, n + 1);
    for (int i = 0; i < n; i++) V.at(i, i) = P::from_int(-1);
    return V;
}

// One SIX::maxm / SIX::minm call split around the GPU part, so that many of them
// can share one batched TwoStageMethod launch:
//   prepare()  verify + normalize (+ explicit dual for min)       [host]
//   -> lp_leq / lp_tgtf go to the two-stage solver                [GPU]
//   finish()   calcFinalSolution                                  [host]
// maxm: lpsol.h:1992-2033.  minm: lpsol.h:1661-1732 + calcDualMaxm :1585-1655.
template <class P>
struct SixJob {
    typedef typename P::T T;
    bool is_min = false;
    Mat<P> tgtf_orig;
    Normalized<P> N;
    int m_rhs = 0, dn = 0, dm = 0;
    Mat<P> lp_leq, lp_tgtf; // the normalised LP handed to TwoStageMethod

    int prepare(bool minimise, const Mat<P> &tgtf, const Mat<P> &vc, const Mat<P> &eq,
                const Mat<P> &leq)
    {
        is_min = minimise;
        tgtf_orig = tgtf;
        const int maxc = !eq.empty() ? eq.c : leq.c; // verify, :1516-1558
        m_rhs = maxc - 1;
        int st = normalize(N, vc, eq, leq, tgtf, m_rhs);
        if (st) return st;
        if (!is_min) { // (a non-standard vc only changes the final feasibility verdict: the caller routes it)
            lp_leq = N.leq;
            lp_tgtf = N.tgtf;
            return 0;
        }
        const int nd_rhs = N.rhs_idx, rows = N.leq.r;
        dn = rows; // dual: one variable per primal row (:1602-1629)
        dm = nd_rhs;
        lp_leq = Mat<P>(dm, dn + 1);
        lp_tgtf = Mat<P>(1, dn + 1);
        for (int i = 0; i < dm; i++)
            for (int j = 0; j < dn; j++) lp_leq.at(i, j) = N.leq.at(j, i);
        lp_leq.mul_all(P::from_int(-1)); // :1607 (also negates the grown zero column)
        for (int i = 0; i < dm; i++) lp_leq.at(i, dn) = N.tgtf.at(0, i);
        for (int j = 0; j < dn; j++) lp_tgtf.at(0, j) = N.leq.at(j, nd_rhs);
        lp_tgtf.mul_all(P::from_int(-1)); // :1619
        return 0;
    }

    // Returns the SIX status; v / sol are written on XP_SIX_SUCC (v = 0 otherwise).
    int finish(const TwoStageResult<P> &R, T &v, std::vector<T> &sol) const
    {
        v = P::zero();
        if (R.status != XP_SIX_SUCC) return R.status;
        if (!is_min) {
            std::vector<T> ss = R.slack_sol;
            ss.resize((size_t)N.rhs_idx + N.leq.r + 2, P::zero());
            final_solution<P>(sol, v, ss, N.vcmap, tgtf_orig, m_rhs);
        } else {
            // y_i = -(coefficient of dual slack i in the final dual objective row), :1713-1716
            std::vector<T> tmp((size_t)dm + 1 + 2 * m_rhs + 4, P::zero());
            for (int k = 0; k < dm; k++) tmp[k] = P::neg(R.tgtf[dn + k]);
            final_solution<P>(sol, v, tmp, N.vcmap, tgtf_orig, m_rhs);
        }
        v = P::reduce(v);
        if (!P::ok()) return XP_ERR_OVERFLOW;
        return XP_SIX_SUCC;
    }
};

// ------------------------------------------------------------------------ MIP
// MIP<Mat,T>::RecusivePart (lpsol.h:2426-2612) unrolled into a resumable state
// machine: next() yields the node LP to solve, feed() consumes its result.  A
// batch of trees can therefore advance in lockstep, one GPU call per wave,
// while every accept / prune decision is replayed per tree in the reference's
// DFS order (fork_count and the incumbent are order dependent, :2474-2497).
template <class P>
struct MipTree {
    typedef typename P::T T;
    struct Frame {
        Mat<P> leq, eq;   // this node's constraints
        int stage = 0;    // 0: solve node, 1: floor child returned, 2: ceil child returned
        int col = 0, sol_ceil = 0;
        bool have_tmp = false;
        T tmpv;
        std::vector<T> tmp_sol;
    };
    Mat<P> tgtf, vc;
    bool is_max = true, is_bin = false;
    int m_rhs = 0, n1 = 0;
    std::vector<Frame> stack;
    std::vector<int32_t> fork_count;
    bool has_best = false;
    T best_v;
    int nodes = 0;
    // result of the most recently finished frame (what RecusivePart returned)
    int ret_status = 0;
    T v;
    std::vector<T> sol;
    bool done = false;

    void start(const Mat<P> &tg, const Mat<P> &vcm, const Mat<P> &eq, const Mat<P> &leq, bool maxm,
               bool bin)
    {
        tgtf = tg;
        vc = vcm;
        is_max = maxm;
        is_bin = bin;
        const int maxc = !eq.empty() ? eq.c : leq.c;
        m_rhs = maxc - 1;
        n1 = tg.c;
        fork_count.assign(n1 + 1, 0);
        best_v = P::zero();
        v = P::zero();
        sol.assign(n1, P::zero());
        Frame f;
        f.leq = leq;
        f.eq = eq;
        stack.clear();
        stack.push_back(f);
        done = false;
    }
    // The LP of the frame on top of the stack (valid while !done and stage == 0).
    const Frame &top() const { return stack.back(); }

    static bool better_than(bool is_max, T cur_best, T cand)
    { // m_cur_best_v < v (max) / > v (min)
        return is_max ? P::lt(cur_best, cand) : P::gt(cur_best, cand);
    }
    void note_best()
    {
        if (!has_best || better_than(is_max, best_v, v)) {
            has_best = true;
            best_v = v;
        }
    }
    std::vector<uint8_t> allow_rational; // rational_indicator (1 x n1, :2626-2630); empty = none
    bool satisfying(int &col)
    { // MIP::is_satisfying, :2363-2408
        if (!allow_rational.empty()) { // entries marked true may stay rational (:2369-2391)
            for (int j = 0; j < n1; j++) {
                if (allow_rational[j]) continue;
                if (!P::is_int(sol[j]) ||
                    (is_bin && !P::eq(sol[j], P::zero()) && !P::eq(sol[j], P::from_int(1)))) {
                    col = j;
                    return false;
                }
            }
            return true;
        }
        for (int j = 0; j < n1; j++) {
            if (is_bin) {
                if (!P::eq(sol[j], P::zero()) && !P::eq(sol[j], P::from_int(1))) {
                    col = j;
                    return false;
                }
            } else if (!P::is_int(sol[j])) {
                col = j;
                return false;
            }
        }
        return true;
    }
    void push_child(const Frame &parent, int col, int rhs_value, bool ceil_side)
    {
        Frame c;
        c.leq = parent.leq;
        c.eq = parent.eq;
        if (is_bin) { // append x_col = value to eq, :2506-2512 / :2548-2553
            c.eq.grow_rows(1, n1);
            c.eq.at(c.eq.r - 1, col) = P::from_int(1);
            c.eq.at(c.eq.r - 1, m_rhs) = P::from_int(rhs_value);
        } else {      // x_col <= floor  or  -x_col <= -ceil, :2514-2520 / :2555-2559
            c.leq.grow_rows(1, n1);
            c.leq.at(c.leq.r - 1, col) = P::from_int(ceil_side ? -1 : 1);
            c.leq.at(c.leq.r - 1, m_rhs) = P::from_int(ceil_side ? -rhs_value : rhs_value);
        }
        stack.push_back(c);
    }
    // A frame finished with `status` (v / sol hold its outputs): unwind.
    void finish(int status)
    {
        for (;;) {
            stack.pop_back();
            ret_status = status;
            if (stack.empty()) {
                done = true;
                return;
            }
            Frame &f = stack.back();
            if (f.stage == 1) { // the floor child just returned, :2527-2543
                if (status < 0) continue; // hard error: propagate
                if (status == XP_IP_SUCC) {
                    f.tmp_sol = sol;
                    f.tmpv = v;
                    f.have_tmp = true;
                    note_best();
                }
                f.stage = 2;
                Frame copy = f;
                push_child(copy, f.col, f.sol_ceil, true); // the ceil child always runs, :2563
                return;
            }
            // f.stage == 2: the ceil child returned, :2563-2611
            if (status < 0) continue;
            if (status == XP_IP_SUCC) {
                if (f.have_tmp && (is_max ? P::gt(f.tmpv, v) : P::lt(f.tmpv, v))) {
                    v = f.tmpv;
                    sol = f.tmp_sol;
                }
                note_best();
            } else if (f.have_tmp) {
                v = f.tmpv;
                sol = f.tmp_sol;
                note_best();
                status = XP_IP_SUCC;
            }
            // loop: this frame now returns `status` to its own parent
        }
    }
    // Consume the LP result of the top frame (stage 0).
    void feed(int six_status, T lp_v, const std::vector<T> &lp_sol)
    {
        nodes++;
        Frame &f = stack.back();
        if (six_status < 0) return finish(six_status);
        v = lp_v;
        if (six_status != XP_SIX_SUCC) { // :2450-2466
            return finish(six_status == XP_SIX_UNBOUND ? XP_IP_UNBOUND : XP_IP_NO_PRI_FEASIBLE_SOL);
        }
        sol = lp_sol;
        int col = 0;
        if (satisfying(col)) return finish(XP_IP_SUCC);
        if (has_best && (is_max ? P::le(v, best_v) : P::ge(v, best_v)))
            return finish(XP_IP_NO_BETTER_THAN_BEST_SOL); // :2474-2485
        if (fork_count[col] >= 1) return finish(XP_IP_NO_PRI_FEASIBLE_SOL); // :2486-2496
        fork_count[col]++;
        int sol_floor = is_bin ? 0 : P::trunc(sol[col]);
        f.col = col;
        f.sol_ceil = is_bin ? 1 : sol_floor + 1;
        f.stage = 1;
        Frame copy = f;
        push_child(copy, col, sol_floor, false);
    }
};

} // namespace xph
