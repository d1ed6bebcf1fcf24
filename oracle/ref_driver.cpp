// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Flat extern "C" shim over the UNMODIFIED reference solver (stevenknown/xpoly).
// It is compiled by oracle/Makefile against the reference sources where they
// lie (/root/reference/src/com/*.cpp, never copied into this repo) into
// oracle/_ref/libxpoly_ref.so.  It is used (a) to pin oracle/xp_oracle.c,
// (b) to generate tests/golden/*.json (tests/golden/make_golden.py), and
// (c) optionally as bench.py's `--impl reference` CPU arm.
//
// Include order follows the reference's own linsys.cpp:28-41 (example.cpp's
// order does not compile, SURVEY.md section 8c caveat 1).  Matrices are filled
// with set() because the variadic sete() is broken on x86-64 (caveat 2).
#include "ltype.h"
#include "comf.h"
#include "strbuf.h"
#include "smempool.h"
#include "rational.h"
#include "flty.h"
#include "sstl.h"
#include "matt.h"
#include "bs.h"
#include "sbs.h"
#include "sgraph.h"
#include "xmat.h"
#include "linsys.h"
#include "lpsol.h"

#include <stdint.h>
#include <string.h>
#include <time.h>

namespace xcom {
extern LONGLONG g_appro_count;  // rational.cpp:188 (non-static)
// MIP<FloatMat,Float> does not compile as shipped (lpsol.h:2245 calls
// Float::format(StrBuf&)); specialise the debug-only dump hook away
// (SURVEY.md section 8c caveat 3).  Does not touch solver behaviour.
template <> bool MIP<FloatMat, Float>::dump_end_six(UINT, Float, FloatMat &) { return true; }
}  // namespace xcom

using namespace xcom;

namespace {

double g_last_seconds = 0.0;
double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void fill_f64(FloatMat & M, int r, int c, const double * src)
{
    M.reinit(r, c);
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) M.set(i, j, Float(src[(size_t)i * c + j]));
}

void fill_rat(RMat & M, int r, int c, const int32_t * src)
{
    M.reinit(r, c);
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) {
            const int32_t * p = src + 2 * ((size_t)i * c + j);
            M.set(i, j, Rational(p[0], p[1]));
        }
}

void neg_identity_f64(FloatMat & vc, int n)
{
    vc.reinit(n, n + 1);
    for (int i = 0; i < n; i++) vc.set(i, i, Float(-1.0));
}

void neg_identity_rat(RMat & vc, int n)
{
    vc.reinit(n, n + 1);
    for (int i = 0; i < n; i++) vc.set(i, i, Rational(-1));
}

void dump_f64(const FloatMat & M, double * dst, size_t cap)
{
    size_t k = 0;
    for (UINT i = 0; i < M.get_row_size(); i++)
        for (UINT j = 0; j < M.get_col_size(); j++) {
            if (k >= cap) return;
            dst[k++] = M.get(i, j).f();
        }
}

void dump_rat(const RMat & M, int32_t * dst, size_t cap)
{
    size_t k = 0;
    for (UINT i = 0; i < M.get_row_size(); i++)
        for (UINT j = 0; j < M.get_col_size(); j++) {
            if (k >= cap) return;
            Rational r = M.get(i, j);
            dst[2 * k] = r.num();
            dst[2 * k + 1] = r.den();
            k++;
        }
}

}  // namespace

extern "C" {

long long ref_appro_count(void) { return (long long)g_appro_count; }

// Wall time of the last ref_two_stage_f64's SIX::TwoStageMethod call alone.
double ref_last_two_stage_seconds(void) { return g_last_seconds; }

// SIX<FloatMat,Float>::maxm / minm  (lpsol.h:1992 / :1661).
// leq m x (n+1), tgtf 1 x (n+1), vc n x (n+1) or NULL (= -I | 0), eq k x (n+1) or k=0.
// sol must hold n+1 doubles.  Returns SIX status.
int ref_six_solve_f64(int is_min, int m, int n, const double * leq, const double * tgtf,
                      const double * vc, int k, const double * eq, unsigned max_iter,
                      double * v, double * sol)
{
    FloatMat L, T, V, E, S;
    fill_f64(L, m, n + 1, leq);
    fill_f64(T, 1, n + 1, tgtf);
    if (vc) fill_f64(V, n, n + 1, vc); else neg_identity_f64(V, n);
    if (k > 0) fill_f64(E, k, n + 1, eq);
    SIX<FloatMat, Float> six;
    six.set_param(0, max_iter);
    Float val;
    UINT st = is_min ? six.minm(val, S, T, V, E, L) : six.maxm(val, S, T, V, E, L);
    *v = val.f();
    if (st == SIX_SUCC) dump_f64(S, sol, (size_t)n + 1);
    return (int)st;
}

int ref_six_solve_rat(int is_min, int m, int n, const int32_t * leq, const int32_t * tgtf,
                      const int32_t * vc, int k, const int32_t * eq, unsigned max_iter,
                      int32_t * v, int32_t * sol)
{
    RMat L, T, V, E, S;
    fill_rat(L, m, n + 1, leq);
    fill_rat(T, 1, n + 1, tgtf);
    if (vc) fill_rat(V, n, n + 1, vc); else neg_identity_rat(V, n);
    if (k > 0) fill_rat(E, k, n + 1, eq);
    SIX<RMat, Rational> six;
    six.set_param(0, max_iter);
    Rational val;
    UINT st = is_min ? six.minm(val, S, T, V, E, L) : six.maxm(val, S, T, V, E, L);
    v[0] = val.num();
    v[1] = val.den();
    if (st == SIX_SUCC) dump_rat(S, sol, (size_t)n + 1);
    return (int)st;
}

// SIX::TwoStageMethod (lpsol.h:1906) on already-normalised input (x >= 0, no
// equalities): the state-level parity hook of SURVEY.md section 8c.  With
// set_param(0,K) it returns SIX_TIME_OUT plus the exact tableau, objective row
// and basis maps after K pivots of the final solveSlackForm call.
// Output capacities: tab m*(n+m+2), otgtf/slack_sol/nvset/bvset/bv2eq n+m+2, eq2bv m.
// dims[0..3] = rows, cols, rhs_idx, slack_sol cols.
int ref_two_stage_f64(int m, int n, const double * leq, const double * tgtf, unsigned max_iter,
                      int * dims, double * tab, double * otgtf, int32_t * eq2bv, int32_t * bv2eq,
                      uint8_t * nvset, uint8_t * bvset, double * maxv, double * slack_sol)
{
    FloatMat L, T, V, S;
    fill_f64(L, m, n + 1, leq);
    fill_f64(T, 1, n + 1, tgtf);
    neg_identity_f64(V, n);
    SIX<FloatMat, Float> six;
    six.set_param(0, max_iter);
    Float val = 0;
    Vector<bool> nv, bv;
    Vector<INT> b2e, e2b;
    INT rhs = n;
    double t0 = now_s();
    UINT st = six.TwoStageMethod(L, V, T, S, val, nv, bv, b2e, e2b, rhs);
    g_last_seconds = now_s() - t0;
    size_t cap = (size_t)n + m + 2;
    dims[0] = L.get_row_size();
    dims[1] = L.get_col_size();
    dims[2] = rhs;
    dims[3] = S.size() ? (int)S.get_col_size() : 0;
    dump_f64(L, tab, (size_t)m * cap);
    dump_f64(T, otgtf, cap);
    if (S.size()) dump_f64(S, slack_sol, cap);
    *maxv = val.f();
    for (int i = 0; i < m; i++) eq2bv[i] = e2b.get(i);
    for (int i = 0; i < (int)cap; i++) {
        bv2eq[i] = (i <= b2e.get_last_idx()) ? b2e.get(i) : -1;
        nvset[i] = (i <= nv.get_last_idx()) ? nv.get(i) : 0;
        bvset[i] = (i <= bv.get_last_idx()) ? bv.get(i) : 0;
    }
    return (int)st;
}

int ref_two_stage_rat(int m, int n, const int32_t * leq, const int32_t * tgtf, unsigned max_iter,
                      int * dims, int32_t * tab, int32_t * otgtf, int32_t * eq2bv, int32_t * bv2eq,
                      uint8_t * nvset, uint8_t * bvset, int32_t * maxv, int32_t * slack_sol)
{
    RMat L, T, V, S;
    fill_rat(L, m, n + 1, leq);
    fill_rat(T, 1, n + 1, tgtf);
    neg_identity_rat(V, n);
    SIX<RMat, Rational> six;
    six.set_param(0, max_iter);
    Rational val = 0;
    Vector<bool> nv, bv;
    Vector<INT> b2e, e2b;
    INT rhs = n;
    UINT st = six.TwoStageMethod(L, V, T, S, val, nv, bv, b2e, e2b, rhs);
    size_t cap = (size_t)n + m + 2;
    dims[0] = L.get_row_size();
    dims[1] = L.get_col_size();
    dims[2] = rhs;
    dims[3] = S.size() ? (int)S.get_col_size() : 0;
    dump_rat(L, tab, (size_t)m * cap);
    dump_rat(T, otgtf, cap);
    if (S.size()) dump_rat(S, slack_sol, cap);
    maxv[0] = val.num();
    maxv[1] = val.den();
    for (int i = 0; i < m; i++) eq2bv[i] = e2b.get(i);
    for (int i = 0; i < (int)cap; i++) {
        bv2eq[i] = (i <= b2e.get_last_idx()) ? b2e.get(i) : -1;
        nvset[i] = (i <= nv.get_last_idx()) ? nv.get(i) : 0;
        bvset[i] = (i <= bv.get_last_idx()) ? bv.get(i) : 0;
    }
    return (int)st;
}

// MIP<Mat,T>::maxm / minm (lpsol.h:2635 / :2680), general-integer or 0-1.
int ref_mip_solve_rat(int is_min, int is_bin, int m, int n, const int32_t * leq,
                      const int32_t * tgtf, int k, const int32_t * eq, int32_t * v, int32_t * sol)
{
    RMat L, T, V, E, S;
    fill_rat(L, m, n + 1, leq);
    fill_rat(T, 1, n + 1, tgtf);
    neg_identity_rat(V, n);
    if (k > 0) fill_rat(E, k, n + 1, eq);
    MIP<RMat, Rational> mip;
    Rational val;
    UINT st = is_min ? mip.minm(val, S, T, V, E, L, is_bin != 0)
                     : mip.maxm(val, S, T, V, E, L, is_bin != 0);
    v[0] = val.num();
    v[1] = val.den();
    if (st == IP_SUCC) dump_rat(S, sol, (size_t)n + 1);
    return (int)st;
}

// ... with rational_indicator (lpsol.h:2626-2657): n+1 flags.
int ref_mip_solve_rat_ri(int is_min, int is_bin, int m, int n, const int32_t * leq, const int32_t * tgtf, int k,
                         const int32_t * eq, const uint8_t * indicator, int32_t * v, int32_t * sol)
{
    RMat L, T, V, E, S;
    fill_rat(L, m, n + 1, leq);
    fill_rat(T, 1, n + 1, tgtf);
    neg_identity_rat(V, n);
    if (k > 0) fill_rat(E, k, n + 1, eq);
    BMat ri(1, n + 1);
    for (int j = 0; j <= n; j++) ri.set(0, j, indicator[j] != 0);
    MIP<RMat, Rational> mip;
    Rational val;
    UINT st = is_min ? mip.minm(val, S, T, V, E, L, is_bin != 0, &ri) : mip.maxm(val, S, T, V, E, L, is_bin != 0, &ri);
    v[0] = val.num();
    v[1] = val.den();
    if (st == IP_SUCC) dump_rat(S, sol, (size_t)n + 1);
    return (int)st;
}

int ref_mip_solve_f64(int is_min, int is_bin, int m, int n, const double * leq,
                      const double * tgtf, int k, const double * eq, double * v, double * sol)
{
    FloatMat L, T, V, E, S;
    fill_f64(L, m, n + 1, leq);
    fill_f64(T, 1, n + 1, tgtf);
    neg_identity_f64(V, n);
    if (k > 0) fill_f64(E, k, n + 1, eq);
    MIP<FloatMat, Float> mip;
    Float val;
    UINT st = is_min ? mip.minm(val, S, T, V, E, L, is_bin != 0)
                     : mip.maxm(val, S, T, V, E, L, is_bin != 0);
    *v = val.f();
    if (st == IP_SUCC) dump_f64(S, sol, (size_t)n + 1);
    return (int)st;
}

// Lineq::has_solution (linsys.cpp:830): the real caller of the small exact (M)IPs.
int ref_has_solution_rat(int m, int n, const int32_t * leq, int k, const int32_t * eq,
                         int is_int_sol, int is_unique_sol)
{
    RMat L, V, E;
    if (m > 0) fill_rat(L, m, n + 1, leq);
    if (k > 0) fill_rat(E, k, n + 1, eq);
    neg_identity_rat(V, n);
    Lineq lin(NULL);
    return lin.has_solution(L, E, V, n, is_int_sol != 0, is_unique_sol != 0) ? 1 : 0;
}

}  // extern "C"
