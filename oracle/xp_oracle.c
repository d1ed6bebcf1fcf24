/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 * See xp_oracle.h.  Scalar semantics restated from the reference's
 * flty.cpp:41-131 (Float) and rational.cpp:76-397 (Rational); the solver body
 * lives in xp_oracle_six.inc and is instantiated once per element type.
 * Build with -ffp-contract=off (oracle/Makefile) so mul and add round
 * separately, as in the reference's shipped flags.
 */
#include "xp_oracle.h"

#include <limits.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ Float */
#define XO_EPS 0.00000000000000001 /* INFINITESIMAL, flty.h:46 */

static int f_eq(double a, double b)
{ /* operator==, flty.cpp:41-58 */
    if ((a > 0 && b < 0) || (a < 0 && b > 0)) return 0;
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    if ((a == 0.0 && b <= XO_EPS) || (b == 0.0 && a <= XO_EPS)) return 1;
    if (a > b) return (a - b) <= XO_EPS;
    return (b - a) <= XO_EPS;
}
static int f_le(double a, double b) { return a < b || f_eq(a, b); } /* flty.cpp:70 */
static int f_ge(double a, double b) { return a > b || f_eq(a, b); } /* flty.cpp:88 */
static int f_is_int(double f)
{ /* Float::is_int, flty.cpp:182-201 */
    double av = f < 0 ? -f : f;
    long long iv = (long long)av;
    double ifv = (double)iv;
    if ((av - ifv) < XO_EPS) return 1;
    ifv = (double)(iv + 1);
    if ((ifv - av) < XO_EPS) return 1;
    return 0;
}

/* --------------------------------------------------------------- Rational */
static long long g_xo_appro = 0;
long long xo_appro_count(void) { return g_xo_appro; }

static long long ll_gcd(long long x, long long y)
{ /* gcdf, rational.cpp:143-158 */
    if (x < 0) x = -x;
    if (y < 0) y = -y;
    if (x > y) {
        long long t = x;
        x = y;
        y = t;
    }
    while (x) {
        long long t = x;
        x = y % x;
        y = t;
    }
    return y;
}

static void ll_reduce(long long *num, long long *den)
{ /* reduce_ll, rational.cpp:163-184 */
    if (*num == 0) {
        *den = 1;
        return;
    }
    long long g = ll_gcd(*num, *den);
    if (g != 1) {
        *num /= g;
        *den /= g;
    }
    if (*den < 0) {
        *den = -*den;
        *num = -*num;
    }
}

static void ll_appro(long long *num, long long *den)
{ /* appro, rational.cpp:189-226: lossy <=7-digit decimal approximation.  The
   * float is compared against double literals and scaled by an int, exactly
   * as written there. */
    g_xo_appro++;
    float v = (float)(*num) / (float)(*den);
    if (v < 100.0) {
        v = v * 1000000;
        *num = (int)v;
        *den = 1000000;
    } else if (v < 1000.0) {
        v = v * 100000;
        *num = (int)v;
        *den = 100000;
    } else if (v < 100000.0) {
        v = v * 10000;
        *num = (int)v;
        *den = 10000;
    } else if (v < 1000000.0) {
        v = v * 1000;
        *num = (int)v;
        *den = 1000;
    } else if (v < 10000000.0) {
        v = v * 100;
        *num = (int)v;
        *den = 100;
    } else if (v < 100000000.0) {
        v = v * 10;
        *num = (int)v;
        *den = 10;
    } else if (v < 2147483647.0) {
        *num = (int)v;
        *den = 1;
    } else {
        *num = 0;
        *den = 1;
    }
    ll_reduce(num, den);
}

static xo_rat r_make(int num, int den)
{
    xo_rat r;
    r.num = num;
    r.den = den;
    return r;
}

/* shared tail of operator* / + and the general case of operator/,
 * rational.cpp:283-307, :337-360, :375-397 */
static xo_rat r_finish(long long rnum, long long rden)
{
    if (rnum == rden) return r_make(1, 1);
    if (rnum == -rden) return r_make(-1, 1);
    if (rden < 0) {
        rnum = -rnum;
        rden = -rden;
    }
    ll_reduce(&rnum, &rden);
    long long t = rnum >= 0 ? rnum : -rnum;
    if (t >= (long long)(INT_MAX >> 2) || rden >= (long long)(INT_MAX >> 2)) {
        ll_reduce(&t, &rden);
        if (t >= (long long)INT_MAX || rden >= (long long)INT_MAX) ll_appro(&t, &rden);
    }
    return r_make((int)(rnum < 0 ? -t : t), (int)rden);
}

static xo_rat r_mul(xo_rat a, xo_rat b)
{ /* operator*, rational.cpp:273-309 */
    long long rnum = (long long)a.num * (long long)b.num;
    if (rnum == 0) return r_make(0, 1);
    return r_finish(rnum, (long long)a.den * (long long)b.den);
}

static xo_rat r_div(xo_rat a, xo_rat b)
{ /* operator/, rational.cpp:312-360 */
    if (a.num == 0) return r_make(0, 1);
    if (a.num == a.den) return b.num < 0 ? r_make(-b.den, -b.num) : r_make(b.den, b.num);
    return r_finish((long long)a.num * (long long)b.den, (long long)a.den * (long long)b.num);
}

static xo_rat r_add(xo_rat a, xo_rat b)
{ /* operator+, rational.cpp:363-397 */
    long long rnum = (long long)a.num * (long long)b.den + (long long)a.den * (long long)b.num;
    if (rnum == 0) return r_make(0, 1);
    return r_finish(rnum, (long long)a.den * (long long)b.den);
}

static xo_rat r_neg(xo_rat a) { return r_make(-a.num, a.den); } /* rational.h:99-104 */
static xo_rat r_sub(xo_rat a, xo_rat b) { return r_add(a, r_neg(b)); } /* rational.h:95-96 */

static int int_gcd(int x, int y)
{ /* Rational::_gcd, rational.cpp:113-127 */
    if (x < 0) x = -x;
    if (y < 0) y = -y;
    if (x > y) {
        int t = x;
        x = y;
        y = t;
    }
    while (x != 0) {
        int t = x;
        x = y % x;
        y = t;
    }
    return y;
}

static xo_rat r_reduce(xo_rat a)
{ /* Rational::reduce, rational.cpp:76-97 */
    if (a.num == 0) return r_make(0, 1);
    int g = int_gcd(a.num, a.den);
    if (g != 1) {
        a.num /= g;
        a.den /= g;
    }
    if (a.den < 0) {
        a.den = -a.den;
        a.num = -a.num;
    }
    return a;
}

/* cross-multiplied comparisons, rational.cpp:229-270; structural ==, rational.h:80-83 */
static int r_lt(xo_rat a, xo_rat b)
{
    return (long long)a.num * b.den < (long long)a.den * b.num;
}
static int r_le(xo_rat a, xo_rat b)
{
    return (long long)a.num * b.den <= (long long)a.den * b.num;
}
static int r_gt(xo_rat a, xo_rat b)
{
    return (long long)a.num * b.den > (long long)a.den * b.num;
}
static int r_ge(xo_rat a, xo_rat b)
{
    return (long long)a.num * b.den >= (long long)a.den * b.num;
}
static int r_eq(xo_rat a, xo_rat b) { return a.num == b.num && a.den == b.den; }

/* ------------------------------------------------- instantiate: FP64 */
#define T double
#define FN(name) name##_f64
#define T_ZERO 0.0
#define T_INT(i) ((double)(i))
#define T_ADD(a, b) ((a) + (b))
#define T_SUB(a, b) ((a) - (b))
#define T_MUL(a, b) ((a) * (b))
#define T_DIV(a, b) ((a) / (b))
#define T_NEG(a) (-(a))
#define T_GT(a, b) ((a) > (b))
#define T_LT(a, b) ((a) < (b))
#define T_EQ(a, b) f_eq((a), (b))
#define T_LE(a, b) f_le((a), (b))
#define T_GE(a, b) f_ge((a), (b))
#define T_REDUCE(a) (a)                  /* Float::reduce is a no-op, flty.h:93 */
#define T_IS_INT(a) f_is_int(a)
#define T_TRUNC(a) ((int)(a))            /* Float::typecast2int, flty.h:85-88 */
#define T_IS_EXACT_ZERO(a) ((a) == 0.0)
#include "xp_oracle_six.inc"
#undef T
#undef FN
#undef T_ZERO
#undef T_INT
#undef T_ADD
#undef T_SUB
#undef T_MUL
#undef T_DIV
#undef T_NEG
#undef T_GT
#undef T_LT
#undef T_EQ
#undef T_LE
#undef T_GE
#undef T_REDUCE
#undef T_IS_INT
#undef T_TRUNC
#undef T_IS_EXACT_ZERO

/* --------------------------------------------- instantiate: Rational */
#define T xo_rat
#define FN(name) name##_rat
#define T_ZERO r_make(0, 1)
#define T_INT(i) r_make((i), 1)
#define T_ADD(a, b) r_add((a), (b))
#define T_SUB(a, b) r_sub((a), (b))
#define T_MUL(a, b) r_mul((a), (b))
#define T_DIV(a, b) r_div((a), (b))
#define T_NEG(a) r_neg(a)
#define T_GT(a, b) r_gt((a), (b))
#define T_LT(a, b) r_lt((a), (b))
#define T_EQ(a, b) r_eq((a), (b))
#define T_LE(a, b) r_le((a), (b))
#define T_GE(a, b) r_ge((a), (b))
#define T_REDUCE(a) r_reduce(a)
#define T_IS_INT(a) ((a).den == 1)       /* RMat::is_imat, xmat.cpp:603-616 */
#define T_TRUNC(a) ((a).num / (a).den)   /* Rational::typecast2int, rational.h:62 */
#define T_IS_EXACT_ZERO(a) ((a).num == 0)
#include "xp_oracle_six.inc"
#undef T
#undef FN

/* ---------------------------------------------------------- MT19937-64
 * Public algorithm (Matsumoto & Nishimura 2004), used only to regenerate the
 * seeded instances SURVEY.md Appendix A4/A5 were probed on: std::mt19937_64
 * feeding std::uniform_real_distribution<double>(0,1), which in libstdc++
 * is double(x) / 2^64 (clamped below 1). */
void xo_mt64_uniform(uint64_t seed, size_t count, double *out)
{
    enum { NN = 312, MM = 156 };
    static const uint64_t MATRIX_A = 0xB5026F5AA96619E9ULL, UM = 0xFFFFFFFF80000000ULL,
                          LM = 0x7FFFFFFFULL;
    uint64_t mt[NN];
    mt[0] = seed;
    for (int i = 1; i < NN; i++) mt[i] = 6364136223846793005ULL * (mt[i - 1] ^ (mt[i - 1] >> 62)) + (uint64_t)i;
    int mti = NN;
    for (size_t k = 0; k < count; k++) {
        if (mti >= NN) {
            for (int i = 0; i < NN; i++) {
                uint64_t x = (mt[i] & UM) | (mt[(i + 1) % NN] & LM);
                mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ ((x & 1ULL) ? MATRIX_A : 0ULL);
            }
            mti = 0;
        }
        uint64_t x = mt[mti++];
        x ^= (x >> 29) & 0x5555555555555555ULL;
        x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
        x ^= (x << 37) & 0xFFF7EEE000000000ULL;
        x ^= (x >> 43);
        double u = (double)x / 18446744073709551616.0;
        if (u >= 1.0) u = 0.99999999999999988897769753748;
        out[k] = u;
    }
}


/* Position-keyed 64-bit checksum of an FP64 matrix: sum over (i, j) of
 * mix64(bits(a[i][j]) ^ mix64(i * C + j)) mod 2^64, splitmix64 finaliser.  The same key
 * function as the product's xp_lp_f64_checksum (k_checksum), restated here so a test can
 * compare the device tableau with the oracle's at sizes where downloading 1 GiB per
 * checkpoint would dominate (K in {1, 10, 50, 200} at c3). */
static uint64_t xo_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t xo_checksum_f64(const double *a, int rows, int cols)
{
    uint64_t s = 0;
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            uint64_t b;
            memcpy(&b, &a[(size_t)i * cols + j], 8);
            s += xo_mix64(b ^ xo_mix64((uint64_t)((size_t)i * cols + j)));
        }
    return s;
}

static double g_xo_last_seconds = 0.0;
double xo_last_solve_seconds(void) { return g_xo_last_seconds; }
static double xo_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------- C entry points */
#define DEFINE_ENTRIES(SFX, TY)                                                                   \
    int xo_six_solve_##SFX(int is_min, int m, int n, const TY *leq, const TY *tgtf,               \
                           const TY *vc, int k, const TY *eq, uint32_t max_iter, TY *v, TY *sol)  \
    {                                                                                             \
        mat_##SFX L = mat_from_##SFX(m, n + 1, leq), Tg = mat_from_##SFX(1, n + 1, tgtf);         \
        mat_##SFX V = vc ? mat_from_##SFX(n, n + 1, vc) : default_vc_##SFX(n);                    \
        mat_##SFX E = k > 0 ? mat_from_##SFX(k, n + 1, eq) : mat_new_##SFX(0, 0);                 \
        int st = is_min ? minm_##SFX(&Tg, &V, &E, &L, max_iter, v, sol)                           \
                        : maxm_##SFX(&Tg, &V, &E, &L, max_iter, v, sol);                          \
        mat_free_##SFX(&L);                                                                       \
        mat_free_##SFX(&Tg);                                                                      \
        mat_free_##SFX(&V);                                                                       \
        mat_free_##SFX(&E);                                                                       \
        return st;                                                                                \
    }                                                                                             \
    int xo_two_stage_##SFX(int m, int n, const TY *leq, const TY *tgtf, uint32_t max_iter,        \
                           int *dims, TY *tab, TY *otgtf, int32_t *eq2bv, int32_t *bv2eq,         \
                           uint8_t *nvset, uint8_t *bvset, TY *maxv, TY *slack_sol,               \
                           int32_t *pivot_log, int log_cap, int *n_log)                           \
    {                                                                                             \
        int cap = n + m + 2;                                                                      \
        six_##SFX S;                                                                              \
        six_init_##SFX(&S, cap + 2, max_iter);                                                    \
        S.tab = mat_from_##SFX(m, n + 1, leq);                                                    \
        S.tgtf = mat_from_##SFX(1, n + 1, tgtf);                                                  \
        S.rhs_idx = n;                                                                            \
        S.nvc = n;                                                                                \
        {                                                                                         \
            mat_##SFX V = default_vc_##SFX(n);                                                    \
            for (int i = 0; i < n; i++) {                                                         \
                S.vc_diag[i] = V.a[(size_t)i * (n + 1) + i];                                      \
                S.vc_rhs[i] = V.a[(size_t)i * (n + 1) + n];                                       \
            }                                                                                     \
            mat_free_##SFX(&V);                                                                   \
        }                                                                                         \
        S.log = pivot_log;                                                                        \
        S.log_cap = log_cap;                                                                      \
        TY *sol = (TY *)calloc((size_t)cap + 2, sizeof(TY));                                      \
        int st = two_stage_##SFX(&S, maxv, sol);                                                  \
        dims[0] = S.tab.r;                                                                        \
        dims[1] = S.tab.c;                                                                        \
        dims[2] = S.rhs_idx;                                                                      \
        dims[3] = (st == XO_SIX_NO_PRI_FEASIBLE_SOL || st < 0) ? 0 : S.tgtf.c;                    \
        memcpy(tab, S.tab.a, (size_t)S.tab.r * (size_t)S.tab.c * sizeof(TY));                     \
        memcpy(otgtf, S.tgtf.a, (size_t)S.tgtf.c * sizeof(TY));                                   \
        if (dims[3]) memcpy(slack_sol, sol, (size_t)S.tgtf.c * sizeof(TY));                       \
        for (int i = 0; i < m; i++) eq2bv[i] = S.eq2bv[i];                                        \
        for (int i = 0; i < cap; i++) {                                                           \
            int live = i < S.rhs_idx;                                                             \
            bv2eq[i] = live ? S.bv2eq[i] : -1;                                                    \
            nvset[i] = live ? S.nvset[i] : 0;                                                     \
            bvset[i] = live ? S.bvset[i] : 0;                                                     \
        }                                                                                         \
        if (n_log) *n_log = S.log_n;                                                              \
        free(sol);                                                                                \
        six_free_##SFX(&S);                                                                       \
        return st;                                                                                \
    }                                                                                             \
    int xo_slack_##SFX(int m, int C, TY *tab, TY *tgtf, uint8_t *nvset, uint8_t *bvset,           \
                       int32_t *bv2eq, int32_t *eq2bv, const TY *vc_diag, const TY *vc_rhs,       \
                       uint32_t max_iter, TY *maxv, TY *sol, uint32_t *iters, int32_t *pivot_log, \
                       int log_cap, int *n_log)                                                   \
    {                                                                                             \
        int rhs = C - 1;                                                                          \
        six_##SFX S;                                                                              \
        six_init_##SFX(&S, C + m + 2, max_iter);                                                  \
        S.tab = mat_from_##SFX(m, C, tab);                                                        \
        S.tgtf = mat_from_##SFX(1, C, tgtf);                                                      \
        S.rhs_idx = rhs;                                                                          \
        S.nvc = rhs;                                                                              \
        {                                                                                         \
            mat_##SFX V = default_vc_##SFX(1);                                                    \
            for (int i = 0; i < rhs; i++) {                                                       \
                S.vc_diag[i] = vc_diag ? vc_diag[i] : V.a[0];                                     \
                S.vc_rhs[i] = vc_rhs ? vc_rhs[i] : V.a[1];                                        \
            }                                                                                     \
            mat_free_##SFX(&V);                                                                   \
        }                                                                                         \
        memcpy(S.nvset, nvset, (size_t)rhs);                                                      \
        memcpy(S.bvset, bvset, (size_t)rhs);                                                      \
        memcpy(S.bv2eq, bv2eq, (size_t)rhs * sizeof(int32_t));                                    \
        memcpy(S.eq2bv, eq2bv, (size_t)m * sizeof(int32_t));                                      \
        S.log = pivot_log;                                                                        \
        S.log_cap = log_cap;                                                                      \
        S.phase = XO_PH_MAIN;                                                                     \
        double t0__ = xo_now();                                                                   \
        int st = solve_slack_##SFX(&S, maxv, sol);                                                \
        g_xo_last_seconds = xo_now() - t0__;                                                      \
        memcpy(tab, S.tab.a, (size_t)m * (size_t)C * sizeof(TY));                                 \
        memcpy(tgtf, S.tgtf.a, (size_t)C * sizeof(TY));                                           \
        memcpy(nvset, S.nvset, (size_t)rhs);                                                      \
        memcpy(bvset, S.bvset, (size_t)rhs);                                                      \
        memcpy(bv2eq, S.bv2eq, (size_t)rhs * sizeof(int32_t));                                    \
        memcpy(eq2bv, S.eq2bv, (size_t)m * sizeof(int32_t));                                      \
        if (iters) *iters = S.last_iters;                                                         \
        if (n_log) *n_log = S.log_n;                                                              \
        six_free_##SFX(&S);                                                                       \
        return st;                                                                                \
    }                                                                                             \
    int xo_mip_solve_##SFX(int is_min, int is_bin, int m, int n, const TY *leq, const TY *tgtf,   \
                           int k, const TY *eq, TY *v, TY *sol, int *n_nodes)                     \
    {                                                                                             \
        mat_##SFX L = mat_from_##SFX(m, n + 1, leq), Tg = mat_from_##SFX(1, n + 1, tgtf);         \
        mat_##SFX V = default_vc_##SFX(n);                                                        \
        mat_##SFX E = k > 0 ? mat_from_##SFX(k, n + 1, eq) : mat_new_##SFX(0, 0);                 \
        int st = mip_solve_##SFX(is_min, is_bin, &Tg, &V, &E, &L, v, sol, n_nodes, NULL);               \
        mat_free_##SFX(&L);                                                                       \
        mat_free_##SFX(&Tg);                                                                      \
        mat_free_##SFX(&V);                                                                       \
        mat_free_##SFX(&E);                                                                       \
        return st;                                                                                \
    }

DEFINE_ENTRIES(f64, double)
DEFINE_ENTRIES(rat, xo_rat)

/* Lineq::has_solution, linsys.cpp:830-906, with SIX::reviseTargetFunc
 * (lpsol.h:2052-2074) applied to the all-ones objective. */
int xo_has_solution_rat(int m, int n, const xo_rat *leq, int k, const xo_rat *eq, int is_int_sol,
                        int is_unique_sol)
{
    if (m * (n + 1) == 0 && k * (n + 1) == 0) return 0;
    mat_rat L = m > 0 ? mat_from_rat(m, n + 1, leq) : mat_new_rat(0, 0);
    mat_rat E = k > 0 ? mat_from_rat(k, n + 1, eq) : mat_new_rat(0, 0);
    mat_rat V = default_vc_rat(n);
    int cols = m > 0 ? n + 1 : 0; /* tgtf(1, leq.get_col_size()), linsys.cpp:851 */
    mat_rat Tg = mat_new_rat(1, cols > 0 ? cols : 1);
    for (int j = 0; j < n && j < Tg.c; j++) {
        int nz = 0;
        if (L.c > 0 && !col_all_eq_rat(&L, j, r_make(0, 1))) nz = 1;
        if (E.c > 0 && !col_all_eq_rat(&E, j, r_make(0, 1))) nz = 1;
        Tg.a[j] = nz ? r_make(1, 1) : r_make(0, 1);
    }
    xo_rat v;
    xo_rat *sol = (xo_rat *)calloc((size_t)n + 2, sizeof(xo_rat));
    int res = 0, st;
    if (is_int_sol) {
        st = mip_solve_rat(0, 0, &Tg, &V, &E, &L, &v, sol, NULL, NULL);
        if (st == XO_IP_SUCC || (!is_unique_sol && st == XO_IP_UNBOUND)) res = 1;
        if (!res) {
            st = mip_solve_rat(1, 0, &Tg, &V, &E, &L, &v, sol, NULL, NULL);
            if (st == XO_IP_SUCC || (!is_unique_sol && st == XO_IP_UNBOUND)) res = 1;
        }
    } else {
        st = maxm_rat(&Tg, &V, &E, &L, 0xFFFFFFFFu, &v, sol);
        if (st == XO_SIX_SUCC || (!is_unique_sol && st == XO_SIX_UNBOUND)) res = 1;
        if (!res) {
            st = minm_rat(&Tg, &V, &E, &L, 0xFFFFFFFFu, &v, sol);
            if (st == XO_SIX_SUCC || (!is_unique_sol && st == XO_SIX_UNBOUND)) res = 1;
        }
    }
    free(sol);
    mat_free_rat(&L);
    mat_free_rat(&E);
    mat_free_rat(&V);
    mat_free_rat(&Tg);
    return res;
}

/* bench.py's CPU baseline for the batched configuration: TwoStageMethod on `batch` LPs of one
 * shape, one after the other (the reference has no batching; callers loop).  Re-entrant, so
 * the bench can run one call per host thread on disjoint slices.  Returns seconds spent. */
double xo_two_stage_f64_many(int batch, int m, int n, const double *leq, const double *tgtf,
                             int32_t *status, double *maxv_out)
{
    int cap = n + m + 2;
    double *tab = (double *)malloc((size_t)m * cap * sizeof(double));
    double *otg = (double *)malloc((size_t)cap * sizeof(double));
    double *ssol = (double *)malloc((size_t)cap * sizeof(double));
    int32_t *eq2bv = (int32_t *)malloc((size_t)m * sizeof(int32_t));
    int32_t *bv2eq = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
    uint8_t *nv = (uint8_t *)malloc((size_t)cap), *bv = (uint8_t *)malloc((size_t)cap);
    double t0 = xo_now();
    for (int k = 0; k < batch; k++) {
        int dims[4];
        double maxv;
        status[k] = xo_two_stage_f64(m, n, leq + (size_t)k * m * (n + 1), tgtf + (size_t)k * (n + 1),
                                     0xFFFFFFFFu, dims, tab, otg, eq2bv, bv2eq, nv, bv, &maxv, ssol,
                                     NULL, 0, NULL);
        if (maxv_out) maxv_out[k] = status[k] == XO_SIX_SUCC ? maxv : 0.0;
    }
    double dt = xo_now() - t0;
    free(tab); free(otg); free(ssol); free(eq2bv); free(bv2eq); free(nv); free(bv);
    return dt;
}

/* The same for the exact side (bench.py's all-core baselines of c4 / c5 / has_solution): plain
 * loops over independent problems, re-entrant (one call per host thread on disjoint slices; the
 * appro counter is a diagnostic and may race).  appro_flag[k] (optional) = 1 when the
 * reference's lossy appro() fired while LP k was being solved -- only meaningful single-threaded. */
double xo_two_stage_rat_many(int batch, int m, int n, const xo_rat *leq, const xo_rat *tgtf,
                             int32_t *status, xo_rat *maxv_out, uint8_t *appro_flag)
{
    int cap = n + m + 2;
    xo_rat *tab = (xo_rat *)malloc((size_t)m * cap * sizeof(xo_rat));
    xo_rat *otg = (xo_rat *)malloc((size_t)cap * sizeof(xo_rat));
    xo_rat *ssol = (xo_rat *)malloc((size_t)cap * sizeof(xo_rat));
    int32_t *eq2bv = (int32_t *)malloc((size_t)m * sizeof(int32_t));
    int32_t *bv2eq = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
    uint8_t *nv = (uint8_t *)malloc((size_t)cap), *bv = (uint8_t *)malloc((size_t)cap);
    double t0 = xo_now();
    for (int k = 0; k < batch; k++) {
        int dims[4];
        xo_rat maxv = {0, 1};
        long long a0 = g_xo_appro;
        status[k] = xo_two_stage_rat(m, n, leq + (size_t)k * m * (n + 1), tgtf + (size_t)k * (n + 1),
                                     0xFFFFFFFFu, dims, tab, otg, eq2bv, bv2eq, nv, bv, &maxv, ssol,
                                     NULL, 0, NULL);
        if (maxv_out) maxv_out[k] = maxv;
        if (appro_flag) appro_flag[k] = g_xo_appro != a0;
    }
    double dt = xo_now() - t0;
    free(tab); free(otg); free(ssol); free(eq2bv); free(bv2eq); free(nv); free(bv);
    return dt;
}

double xo_mip_solve_rat_many(int batch, int is_min, int is_bin, int m, int n, const xo_rat *leq,
                             const xo_rat *tgtf, int32_t *status, xo_rat *v_out, int32_t *nodes)
{
    xo_rat *sol = (xo_rat *)malloc((size_t)(n + 1) * sizeof(xo_rat));
    double t0 = xo_now();
    for (int k = 0; k < batch; k++) {
        xo_rat v = {0, 1};
        int nn = 0;
        status[k] = xo_mip_solve_rat(is_min, is_bin, m, n, leq + (size_t)k * m * (n + 1),
                                     tgtf + (size_t)k * (n + 1), 0, NULL, &v, sol, &nn);
        if (v_out) v_out[k] = v;
        if (nodes) nodes[k] = nn;
    }
    double dt = xo_now() - t0;
    free(sol);
    return dt;
}

/* systems of different sizes: system k has ms[k] rows of ns[k]+1 entries at pool + off[k] */
double xo_has_solution_rat_many(int batch, const int32_t *ms, const int32_t *ns, const int64_t *off,
                                const xo_rat *pool, int is_int_sol, int is_unique_sol, int32_t *res)
{
    double t0 = xo_now();
    for (int k = 0; k < batch; k++)
        res[k] = xo_has_solution_rat(ms[k], ns[k], pool + off[k], 0, NULL, is_int_sol, is_unique_sol);
    return xo_now() - t0;
}

/* MIP::maxm / minm with rational_indicator (lpsol.h:2626-2657): n+1 flags, non-zero = that
 * entry of the solution may stay rational. */
int xo_mip_solve_rat_ri(int is_min, int is_bin, int m, int n, const xo_rat *leq, const xo_rat *tgtf, int k,
                        const xo_rat *eq, const uint8_t *rational_indicator, xo_rat *v, xo_rat *sol, int *n_nodes)
{
    mat_rat L = mat_from_rat(m, n + 1, leq), Tg = mat_from_rat(1, n + 1, tgtf), V = default_vc_rat(n);
    mat_rat E = k > 0 ? mat_from_rat(k, n + 1, eq) : mat_new_rat(0, 0);
    int st = mip_solve_rat(is_min, is_bin, &Tg, &V, &E, &L, v, sol, n_nodes, rational_indicator);
    mat_free_rat(&L);
    mat_free_rat(&Tg);
    mat_free_rat(&V);
    mat_free_rat(&E);
    return st;
}
