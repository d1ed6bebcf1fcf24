// Adaptor for the reference tree: routes SIX<FloatMat,Float>, SIX<RMat,Rational> and
// MIP<RMat,Rational> (maxm / minm) of stevenknown/xpoly to libxpoly_b200.so.
//
// Usage (in a translation unit of the reference, after its usual includes -- the order of
// linsys.cpp:28-41 ending in "lpsol.h"):
//     #include "xp_six.hpp"
// The explicit specialisations below replace the generic template bodies of lpsol.h
// (maxm :1992, minm :1661, MIP::maxm :2635, MIP::minm :2680) for these instantiations only;
// SIX<FloatMat,Float>::TwoStageMethod is routed too; every other member (set_param,
// reviseTargetFunc, ...) stays the reference's.  Link with -lxpoly_b200 -lcudart.  See INTEGRATION.md.
//
// This header needs the reference's headers and is therefore NOT compiled into the product
// library; tests/test_adaptor_cpu.py compiles and links it where /root/reference exists.
#pragma once

#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "xpoly_b200.h"

namespace xcom {

struct XpCtxHolder { // one context per host thread (the reference is single-threaded)
    xp_ctx *ctx;
    XpCtxHolder() : ctx(NULL)
    {
        if (xp_ctx_create(0, &ctx) != 0) {
            fprintf(stderr, "xpoly_b200: %s\n", xp_last_error(ctx));
            abort(); // no CPU fallback
        }
    }
    ~XpCtxHolder() { xp_ctx_destroy(ctx); }
};
inline xp_ctx *xp_thread_ctx()
{
    static thread_local XpCtxHolder h;
    return h.ctx;
}

template <class Mat> inline void *xp_raw(Mat const &m)
{ // Matrix<T>::get_matrix() is non-const (matt.h:283); an empty matrix maps to NULL
    Mat &mm = const_cast<Mat &>(m);
    return mm.size() ? (void *)mm.get_matrix() : NULL;
}

inline UINT xp_status(int st)
{
    if (st < 0 && st != XP_ERR_REFERENCE_UB && st != XP_ERR_OVERFLOW) {
        fprintf(stderr, "xpoly_b200: error %d: %s\n", st, xp_last_error(xp_thread_ctx()));
        abort();
    }
    return (UINT)st; // SIX_SUCC .. SIX_TIME_OUT / IP_SUCC .. unchanged (lpsol.h:198-202, :2082-2085)
}

// ---------------------------------------------------------------- SIX<FloatMat,Float>
template <>
inline UINT SIX<FloatMat, Float>::maxm(OUT Float &maxv, OUT FloatMat &sol, FloatMat const &tgtf, IN FloatMat &vc,
                                       FloatMat const &eq, FloatMat const &leq, INT rhs_idx)
{
    const int n = (int)tgtf.get_col_size() - 1; // the last column is the constant (lpsol.h:1516-1558)
    ASSERT(rhs_idx == -1 || rhs_idx == n, ("unsupported rhs_idx"));
    sol.reinit(1, n + 1); // lpsol.h:1880
    double v = 0.0;
    int st = xp_six_maxm_f64(xp_thread_ctx(), (int)leq.get_row_size(), n, (const double *)xp_raw(tgtf),
                             (const double *)xp_raw(vc), (int)eq.get_row_size(), (const double *)xp_raw(eq),
                             (const double *)xp_raw(leq), m_max_iter, &v, (double *)sol.get_matrix(), NULL);
    maxv = Float(v);
    return xp_status(st);
}
template <>
inline UINT SIX<FloatMat, Float>::minm(OUT Float &minv, OUT FloatMat &sol, FloatMat const &tgtf, IN FloatMat &vc,
                                       FloatMat const &eq, FloatMat const &leq, INT rhs_idx)
{
    const int n = (int)tgtf.get_col_size() - 1;
    ASSERT(rhs_idx == -1 || rhs_idx == n, ("unsupported rhs_idx"));
    sol.reinit(1, n + 1);
    double v = 0.0;
    int st = xp_six_minm_f64(xp_thread_ctx(), (int)leq.get_row_size(), n, (const double *)xp_raw(tgtf),
                             (const double *)xp_raw(vc), (int)eq.get_row_size(), (const double *)xp_raw(eq),
                             (const double *)xp_raw(leq), m_max_iter, &v, (double *)sol.get_matrix(), NULL);
    minv = Float(v);
    return xp_status(st);
}

// SIX<FloatMat,Float>::TwoStageMethod (lpsol.h:1906-1930) -- the one public entry that exposes
// the solver state.  IN: the normalised LP (newleq m x (n+1), newtgtf 1 x (n+1), newvc n x (n+1),
// new_rhs_idx = n).  OUT, exactly as stage1 + slack + solveSlackForm leave them (:1405-1433,
// :1783-1844, :1007-1191): newleq = final tableau m x (n+m+1), newtgtf = final objective row,
// newvc grown by the slack variables' -1 diagonal, slack_sol, maxv, the four basis maps and
// new_rhs_idx = n + m.  Phase 1 (auxiliary LP) runs on the device as well.  On
// SIX_NO_PRI_FEASIBLE_SOL the reference returns out of stage1 with its arguments half reshaped;
// here they are left untouched.
template <>
inline UINT SIX<FloatMat, Float>::TwoStageMethod(IN OUT FloatMat &newleq, IN OUT FloatMat &newvc,
                                                 IN OUT FloatMat &newtgtf, IN OUT FloatMat &slack_sol,
                                                 IN OUT Float &maxv, IN OUT Vector<bool> &nvset,
                                                 IN OUT Vector<bool> &bvset, IN OUT Vector<INT> &bv2eqmap,
                                                 IN OUT Vector<INT> &eq2bvmap, IN OUT INT &new_rhs_idx)
{
    const int m = (int)newleq.get_row_size(), n = (int)new_rhs_idx, C = n + m + 1;
    ASSERT((int)newleq.get_col_size() == n + 1 && (int)newtgtf.get_col_size() == n + 1, ("normalised LP expected"));
    std::vector<double> vd(n, -1.0), vr(n, 0.0), ss(C, 0.0), tg(C, 0.0);
    for (int j = 0; j < n && j < (int)newvc.get_row_size(); j++) { // the entries is_feasible reads, :798-802
        vd[j] = newvc.get(j, j).f();
        vr[j] = newvc.get(j, n).f();
    }
    std::vector<int32_t> e2b(m, 0), b2e(C - 1, 0);
    std::vector<uint8_t> nv(C - 1, 0), bv(C - 1, 0);
    int32_t st = 0;
    double mv = 0.0;
    int rc = xp_six_two_stage_f64_large_vc(xp_thread_ctx(), m, n, (const double *)xp_raw(newleq),
                                           (const double *)xp_raw(newtgtf), vd.data(), vr.data(), m_max_iter,
                                           XP_RULE_REFERENCE, &st, &mv, ss.data(), tg.data(), e2b.data(), NULL, NULL);
    if (rc) return xp_status(rc);
    if (st == XP_SIX_NO_PRI_FEASIBLE_SOL) return (UINT)st;
    newvc.insertColumnsBefore(n, m); // SIX::slack, :1414-1432
    newvc.grow_row(m);
    for (int i = 0; i < m; i++) newvc.set(n + i, n + i, Float(-1.0));
    newleq.reinit(m, C);
    rc = xp_ctx_last_lp_download(xp_thread_ctx(), (double *)newleq.get_matrix(), NULL, nv.data(), bv.data(),
                                 b2e.data(), NULL);
    if (rc) return xp_status(rc);
    newtgtf.reinit(1, C);
    slack_sol.reinit(1, C);
    for (int j = 0; j < C; j++) {
        newtgtf.set(0, j, Float(tg[j]));
        slack_sol.set(0, j, Float(ss[j]));
    }
    for (int j = 0; j < C - 1; j++) {
        nvset.set(j, nv[j] != 0);
        bvset.set(j, bv[j] != 0);
        bv2eqmap.set(j, b2e[j]);
    }
    for (int i = 0; i < m; i++) eq2bvmap.set(i, e2b[i]);
    maxv = Float(mv);
    new_rhs_idx = n + m;
    return (UINT)st;
}

// ---------------------------------------------------------------- SIX<RMat,Rational>
// XP_ERR_OVERFLOW: the exact result does not fit int32/int32 -- the case in which the reference
// would have silently replaced it by a 7-digit approximation (rational.cpp:189-226).
template <>
inline UINT SIX<RMat, Rational>::maxm(OUT Rational &maxv, OUT RMat &sol, RMat const &tgtf, IN RMat &vc,
                                      RMat const &eq, RMat const &leq, INT rhs_idx)
{
    const int n = (int)tgtf.get_col_size() - 1;
    ASSERT(rhs_idx == -1 || rhs_idx == n, ("unsupported rhs_idx"));
    sol.reinit(1, n + 1);
    xp_rat v = {0, 1};
    int st = xp_six_maxm_rat(xp_thread_ctx(), (int)leq.get_row_size(), n, (const xp_rat *)xp_raw(tgtf),
                             (const xp_rat *)xp_raw(vc), (int)eq.get_row_size(), (const xp_rat *)xp_raw(eq),
                             (const xp_rat *)xp_raw(leq), m_max_iter, &v, (xp_rat *)sol.get_matrix(), NULL);
    maxv = Rational(v.num, v.den);
    return xp_status(st);
}
template <>
inline UINT SIX<RMat, Rational>::minm(OUT Rational &minv, OUT RMat &sol, RMat const &tgtf, IN RMat &vc,
                                      RMat const &eq, RMat const &leq, INT rhs_idx)
{
    const int n = (int)tgtf.get_col_size() - 1;
    ASSERT(rhs_idx == -1 || rhs_idx == n, ("unsupported rhs_idx"));
    sol.reinit(1, n + 1);
    xp_rat v = {0, 1};
    int st = xp_six_minm_rat(xp_thread_ctx(), (int)leq.get_row_size(), n, (const xp_rat *)xp_raw(tgtf),
                             (const xp_rat *)xp_raw(vc), (int)eq.get_row_size(), (const xp_rat *)xp_raw(eq),
                             (const xp_rat *)xp_raw(leq), m_max_iter, &v, (xp_rat *)sol.get_matrix(), NULL);
    minv = Rational(v.num, v.den);
    return xp_status(st);
}

// BMat is Matrix<bool> (xmat.h:166): one byte per flag, row-major
inline const uint8_t *xp_indicator(BMat *ri)
{
    static_assert(sizeof(bool) == 1, "bool flags are passed as bytes");
    return ri ? (const uint8_t *)ri->get_matrix() : NULL;
}

// ---------------------------------------------------------------- MIP<RMat,Rational>
// vc must be -I | 0 (MIP::verify, lpsol.h:2349-2358), which is what the C ABI assumes.
template <>
inline UINT MIP<RMat, Rational>::maxm(OUT Rational &maxv, OUT RMat &sol, RMat const &tgtf, IN RMat &vc,
                                      RMat const &eq, RMat const &leq, bool is_bin, IN BMat *rational_indicator,
                                      INT rhs_idx)
{
    ASSERT(rational_indicator == NULL || (rational_indicator->get_row_size() == 1 &&
                                          rational_indicator->get_col_size() == tgtf.get_col_size()),
           ("rational_indicator must be 1 x (n+1)")); // MIP::verify, lpsol.h:2341-2346
    const int n = (int)tgtf.get_col_size() - 1;
    ASSERT(rhs_idx == -1 || rhs_idx == n, ("unsupported rhs_idx"));
    (void)vc;
    sol.reinit(1, n + 1);
    xp_rat v = {0, 1};
    int32_t nodes = 0;
    int st = xp_mip_solve_rat_ri(xp_thread_ctx(), /*is_min=*/0, is_bin ? 1 : 0, (int)leq.get_row_size(), n,
                                 (const xp_rat *)xp_raw(tgtf), (int)eq.get_row_size(), (const xp_rat *)xp_raw(eq),
                                 (const xp_rat *)xp_raw(leq), xp_indicator(rational_indicator), &v,
                                 (xp_rat *)sol.get_matrix(), &nodes);
    maxv = Rational(v.num, v.den);
    m_times = (UINT)nodes; // lpsol.h:2443
    return xp_status(st);
}
template <>
inline UINT MIP<RMat, Rational>::minm(OUT Rational &minv, OUT RMat &sol, RMat const &tgtf, IN RMat &vc,
                                      RMat const &eq, RMat const &leq, bool is_bin, IN BMat *rational_indicator,
                                      INT rhs_idx)
{
    ASSERT(rational_indicator == NULL || (rational_indicator->get_row_size() == 1 &&
                                          rational_indicator->get_col_size() == tgtf.get_col_size()),
           ("rational_indicator must be 1 x (n+1)")); // MIP::verify, lpsol.h:2341-2346
    const int n = (int)tgtf.get_col_size() - 1;
    ASSERT(rhs_idx == -1 || rhs_idx == n, ("unsupported rhs_idx"));
    (void)vc;
    sol.reinit(1, n + 1);
    xp_rat v = {0, 1};
    int32_t nodes = 0;
    int st = xp_mip_solve_rat_ri(xp_thread_ctx(), /*is_min=*/1, is_bin ? 1 : 0, (int)leq.get_row_size(), n,
                                 (const xp_rat *)xp_raw(tgtf), (int)eq.get_row_size(), (const xp_rat *)xp_raw(eq),
                                 (const xp_rat *)xp_raw(leq), xp_indicator(rational_indicator), &v,
                                 (xp_rat *)sol.get_matrix(), &nodes);
    minv = Rational(v.num, v.den);
    m_times = (UINT)nodes;
    return xp_status(st);
}

// ---------------------------------------------------------------- Lineq::has_solution, batched
// The producer side (DepPolyMgr::buildDepPoly, poly.cpp:1166-1195) asks one has_solution
// question per reference pair and loop depth, each synchronously (linsys.cpp:830-906).  This
// collector keeps the call shape -- add() takes exactly has_solution's arguments -- and answers
// all collected systems with one xp_has_solution_rat_ragged call.  vc must be the default -I | 0
// (what every caller in the reference passes).
class XpHasSolutionBatch {
    std::vector<int32_t> m_ns, m_ms, m_ks, m_res;
    std::vector<int64_t> m_lo, m_eo;
    std::vector<xp_rat> m_lp, m_ep;
    UINT m_failed = 0;
    static void append(std::vector<xp_rat> &pool, RMat const &m)
    {
        const xp_rat *p = (const xp_rat *)xp_raw(m);
        if (p) pool.insert(pool.end(), p, p + (size_t)m.get_row_size() * m.get_col_size());
    }

public:
    // Returns the index of the query.
    UINT add(RMat const &leq, RMat const &eq, UINT rhs_idx)
    {
        ASSERT0(leq.size() == 0 || leq.get_col_size() == rhs_idx + 1);
        ASSERT0(eq.size() == 0 || eq.get_col_size() == rhs_idx + 1);
        m_ns.push_back((int32_t)rhs_idx);
        m_ms.push_back((int32_t)(leq.size() ? leq.get_row_size() : 0));
        m_ks.push_back((int32_t)(eq.size() ? eq.get_row_size() : 0));
        m_lo.push_back((int64_t)m_lp.size());
        m_eo.push_back((int64_t)m_ep.size());
        append(m_lp, leq);
        append(m_ep, eq);
        return (UINT)m_ns.size() - 1;
    }
    UINT size() const { return (UINT)m_ns.size(); }
    // Answers every collected query (is_int_sol / is_unique_sol as in linsys.cpp:830).
    void run(bool is_int_sol, bool is_unique_sol)
    {
        m_res.assign(m_ns.size(), 0);
        if (m_ns.empty()) return;
        xp_rat dummy = {0, 1};
        int st = xp_has_solution_rat_ragged(xp_thread_ctx(), (int)m_ns.size(), m_ns.data(), m_ms.data(),
                                            m_lo.data(), m_lp.empty() ? &dummy : m_lp.data(), m_lp.size(),
                                            m_ks.data(), m_eo.data(), m_ep.empty() ? &dummy : m_ep.data(),
                                            m_ep.size(), is_int_sol ? 1 : 0, is_unique_sol ? 1 : 0,
                                            m_res.data());
        xp_status(st);
        m_failed = 0;
        for (size_t i = 0; i < m_res.size(); i++) m_failed += m_res[i] < 0;
    }
    // has_solution's answer for query i.  A query the solver could not decide (negative code:
    // XP_ERR_OVERFLOW past int64, XP_ERR_REFERENCE_UB where the reference itself has undefined
    // behaviour, XP_ERR_BAD_ARG) answers TRUE: in the producer (DepPoly::is_empty,
    // poly.cpp:530-573) "no solution" drops the dependence, so an undecided system must stay a
    // dependence -- the conservative direction for a compiler.  failed() / raw(i) tell a caller
    // which answers were defaulted, e.g. to re-ask the synchronous Lineq::has_solution.
    bool get(UINT i) const { return m_res[i] != 0; }
    INT raw(UINT i) const { return m_res[i]; }
    UINT failed() const { return m_failed; }
    void clean()
    {
        m_ns.clear(); m_ms.clear(); m_ks.clear(); m_res.clear();
        m_lo.clear(); m_eo.clear(); m_lp.clear(); m_ep.clear();
    }
};

} // namespace xcom
