/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * xp_oracle: a plain-C, single-threaded CPU restatement of the simplex hot path
 * of stevenknown/xpoly (SIX / MIP over Float and Rational), written from the
 * algorithm, on flat arrays.  Every function cites the reference file:line it
 * follows.  It is the checker for the CUDA path in tests/, in
 * __graft_entry__.smoke() and in bench.py's cpu_baseline leg -- nothing else
 * may import, link or call it.
 *
 * PINNING: tests/test_oracle_vs_golden.py checks this oracle against
 * the tests/golden/ JSON fixtures, which were produced by running the UNMODIFIED reference
 * (oracle/_ref/libxpoly_ref.so, built by oracle/Makefile from /root/reference)
 * through tests/golden/make_golden.py, and -- when oracle/_ref is present --
 * tests/test_oracle_vs_ref.py differential-tests it live on seeded LPs.
 *
 * Reference citations are relative to /root/reference/src/com/.
 */
#ifndef XP_ORACLE_H
#define XP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: lpsol.h:198-202 and :2082-2085. */
#define XO_SIX_SUCC 0
#define XO_SIX_UNBOUND 1
#define XO_SIX_NO_PRI_FEASIBLE_SOL 2
#define XO_SIX_OPTIMAL_IS_INFEASIBLE 3
#define XO_SIX_TIME_OUT 4
#define XO_IP_SUCC 0
#define XO_IP_UNBOUND 1
#define XO_IP_NO_PRI_FEASIBLE_SOL 2
#define XO_IP_NO_BETTER_THAN_BEST_SOL 3
/* The reference would have executed undefined behaviour (out-of-bounds read or
 * division by zero in convertEq2Ineq, lpsol.h:1232). */
#define XO_ERR_REFERENCE_UB (-100)

/* Pivot-log phases (one int32 quadruple {phase, nv, bv, row} per pivot() call). */
#define XO_PH_AUX_FORCED 0 /* lpsol.h:908  */
#define XO_PH_AUX_LOOP 1   /* lpsol.h:912  */
#define XO_PH_AUX_XA_OUT 2 /* lpsol.h:939  */
#define XO_PH_MAIN 3       /* lpsol.h:1926 */

typedef struct {
    int32_t num, den; /* layout of xcom::Rational, rational.h:51-52 */
} xo_rat;

/* Number of lossy appro() calls so far (rational.cpp:188-226). */
long long xo_appro_count(void);

/* SIX::maxm (lpsol.h:1992) / SIX::minm (lpsol.h:1661).
 * leq m x (n+1), tgtf 1 x (n+1), vc n x (n+1) or NULL (= -I | 0), eq k x (n+1).
 * sol: n+1 entries (valid on XO_SIX_SUCC). */
int xo_six_solve_f64(int is_min, int m, int n, const double *leq, const double *tgtf,
                     const double *vc, int k, const double *eq, uint32_t max_iter, double *v,
                     double *sol);
int xo_six_solve_rat(int is_min, int m, int n, const xo_rat *leq, const xo_rat *tgtf,
                     const xo_rat *vc, int k, const xo_rat *eq, uint32_t max_iter, xo_rat *v,
                     xo_rat *sol);

/* SIX::TwoStageMethod (lpsol.h:1906) with vc = -I.  Output capacities:
 * tab m*(n+m+2); otgtf, slack_sol, nvset, bvset, bv2eq: n+m+2; eq2bv: m.
 * dims = {rows, cols, rhs_idx, slack_sol cols}.  pivot_log (optional) receives
 * up to log_cap quadruples, *n_log the number of pivot() calls made. */
int xo_two_stage_f64(int m, int n, const double *leq, const double *tgtf, uint32_t max_iter,
                     int *dims, double *tab, double *otgtf, int32_t *eq2bv, int32_t *bv2eq,
                     uint8_t *nvset, uint8_t *bvset, double *maxv, double *slack_sol,
                     int32_t *pivot_log, int log_cap, int *n_log);
int xo_two_stage_rat(int m, int n, const xo_rat *leq, const xo_rat *tgtf, uint32_t max_iter,
                     int *dims, xo_rat *tab, xo_rat *otgtf, int32_t *eq2bv, int32_t *bv2eq,
                     uint8_t *nvset, uint8_t *bvset, xo_rat *maxv, xo_rat *slack_sol,
                     int32_t *pivot_log, int log_cap, int *n_log);

/* SIX::solveSlackForm (lpsol.h:1007-1191) alone, in place, on a caller-built
 * slack form: tab m x C (C = rhs_idx+1), tgtf 1 x C, basis maps as in the
 * reference.  vc_diag / vc_rhs (rhs_idx entries each) may be NULL (= -1 / 0).
 * This is the kernel-level oracle for xp_six_slack_*.  sol: C entries. */
int xo_slack_f64(int m, int C, double *tab, double *tgtf, uint8_t *nvset, uint8_t *bvset,
                 int32_t *bv2eq, int32_t *eq2bv, const double *vc_diag, const double *vc_rhs,
                 uint32_t max_iter, double *maxv, double *sol, uint32_t *iters,
                 int32_t *pivot_log, int log_cap, int *n_log);
int xo_slack_rat(int m, int C, xo_rat *tab, xo_rat *tgtf, uint8_t *nvset, uint8_t *bvset,
                 int32_t *bv2eq, int32_t *eq2bv, const xo_rat *vc_diag, const xo_rat *vc_rhs,
                 uint32_t max_iter, xo_rat *maxv, xo_rat *sol, uint32_t *iters,
                 int32_t *pivot_log, int log_cap, int *n_log);

/* MIP::maxm / minm (lpsol.h:2635 / :2680); vc = -I.  *n_nodes = SIX solves made. */
int xo_mip_solve_f64(int is_min, int is_bin, int m, int n, const double *leq, const double *tgtf,
                     int k, const double *eq, double *v, double *sol, int *n_nodes);
int xo_mip_solve_rat(int is_min, int is_bin, int m, int n, const xo_rat *leq, const xo_rat *tgtf,
                     int k, const xo_rat *eq, xo_rat *v, xo_rat *sol, int *n_nodes);

/* ... with MIP's rational_indicator (lpsol.h:2626-2657): n+1 flags, non-zero = may stay rational. */
int xo_mip_solve_rat_ri(int is_min, int is_bin, int m, int n, const xo_rat *leq, const xo_rat *tgtf, int k,
                        const xo_rat *eq, const uint8_t *rational_indicator, xo_rat *v, xo_rat *sol, int *n_nodes);

/* Lineq::has_solution (linsys.cpp:830-906), vc = -I. */
int xo_has_solution_rat(int m, int n, const xo_rat *leq, int k, const xo_rat *eq, int is_int_sol,
                        int is_unique_sol);

/* Wall time of the last xo_slack_* solve loop alone (bench.py cpu_baseline). */
double xo_last_solve_seconds(void);
/* TwoStageMethod on `batch` FP64 LPs of one shape in a loop; returns the seconds spent. */
double xo_two_stage_f64_many(int batch, int m, int n, const double *leq, const double *tgtf,
                             int32_t *status, double *maxv_out /* optional: maxv on SUCC, else 0 */);

/* Exact-side loops for bench.py's CPU baselines (one call per host thread on disjoint slices). */
double xo_two_stage_rat_many(int batch, int m, int n, const xo_rat *leq, const xo_rat *tgtf,
                             int32_t *status, xo_rat *maxv_out, uint8_t *appro_flag);
double xo_mip_solve_rat_many(int batch, int is_min, int is_bin, int m, int n, const xo_rat *leq,
                             const xo_rat *tgtf, int32_t *status, xo_rat *v_out, int32_t *nodes);
double xo_has_solution_rat_many(int batch, const int32_t *ms, const int32_t *ns, const int64_t *off,
                                const xo_rat *pool, int is_int_sol, int is_unique_sol, int32_t *res);

/* Position-keyed checksum of an FP64 matrix (same key function as xp_lp_f64_checksum). */
uint64_t xo_checksum_f64(const double *a, int rows, int cols);

/* std::mt19937_64 + uniform_real_distribution<double>(0,1) stream (libstdc++). */
void xo_mt64_uniform(uint64_t seed, size_t count, double *out);

#ifdef __cplusplus
}
#endif
#endif
