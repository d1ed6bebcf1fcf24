/* xpoly_b200 -- C ABI of the B200-native simplex hot path.
 *
 * Drop-in boundary for the one hot path of stevenknown/xpoly's LP/MIP solvers:
 * pricing -> min-ratio test -> tableau pivot (rank-1 elimination), FP64 and an
 * exact fraction-free integer twin of the Rational simplex.  The reference has
 * no FFI of its own (its API is C++ templates in src/com/lpsol.h); each entry
 * point below names the reference interface it replaces, and
 * xpoly_b200/host/xp_six.hpp (compiled against the reference headers) wraps them with the reference's own signatures
 * (SIX<Mat,T>::maxm/minm/TwoStageMethod/set_param, MIP<Mat,T>::maxm/minm).
 * INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *  - All matrices are dense row-major, element (r,c) at base[r*cols + c]
 *    (the layout of xcom::Matrix<T>::m_mat, matt.h:152-156,289-295), so
 *    FloatMat::get_matrix() / RMat::get_matrix() pointers pass straight in.
 *  - Rationals are {int32 num; int32 den} pairs (xcom::Rational, rational.h:51-52).
 *  - Return value >= 0 is the reference's own status code (lpsol.h:198-202,
 *    :2082-2085); negative values are errors the reference cannot express.
 *  - The library never frees or reallocates caller memory.  Device buffers are
 *    owned by an xp_ctx; one ctx per host thread / CUDA stream.
 *  - There is no CPU fallback: without a usable CUDA device every compute
 *    entry point returns XP_ERR_CUDA.
 */
#ifndef XPOLY_B200_H
#define XPOLY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (lpsol.h:198-202) ---- */
#define XP_SIX_SUCC 0
#define XP_SIX_UNBOUND 1
#define XP_SIX_NO_PRI_FEASIBLE_SOL 2
#define XP_SIX_OPTIMAL_IS_INFEASIBLE 3
#define XP_SIX_TIME_OUT 4
/* ---- MIP status codes (lpsol.h:2082-2085) ---- */
#define XP_IP_SUCC 0
#define XP_IP_UNBOUND 1
#define XP_IP_NO_PRI_FEASIBLE_SOL 2
#define XP_IP_NO_BETTER_THAN_BEST_SOL 3
/* ---- errors (never 0..4) ---- */
#define XP_ERR_CUDA (-1)        /* CUDA runtime / no device; see xp_last_error */
#define XP_ERR_BAD_ARG (-2)     /* dimensions / null pointers */
#define XP_ERR_TOO_LARGE (-3)   /* LP does not fit the batched (shared-memory) path */
#define XP_ERR_OVERFLOW (-4)    /* fraction-free path left int64 / result left int32 */
#define XP_ERR_PEER (-5)        /* sharded LP: a peer GPU did not answer in time */
#define XP_ERR_REFERENCE_UB (-100) /* the reference itself would hit undefined behaviour
                                      (convertEq2Ineq lpsol.h:1232) */

/* Pivot rule.  Only XP_RULE_REFERENCE has an oracle: lowest-index entering
 * variable with c_j > 0 (lpsol.h:1054-1069), first strict minimum ratio row
 * (lpsol.h:603-611), pair-tabu anti-cycling (lpsol.h:68-154). */
#define XP_RULE_REFERENCE 0

#define XP_NO_ITER_LIMIT 0xFFFFFFFFu /* SIX::set_param default, lpsol.h:260 */

typedef struct xp_ctx xp_ctx;
typedef struct {
    int32_t num, den;
} xp_rat;

/* ------------------------------------------------------------------ context */
int xp_ctx_create(int device, xp_ctx **out);
void xp_ctx_destroy(xp_ctx *ctx);
const char *xp_last_error(const xp_ctx *ctx);
const char *xp_version(void);
/* Kernel launches issued by this ctx since creation (for bench accounting). */
uint64_t xp_ctx_launch_count(const xp_ctx *ctx);
/* Device time in ms of the last solve call's kernel region, measured with CUDA
 * events on the ctx stream (excludes H2D/D2H). */
float xp_ctx_last_kernel_ms(const xp_ctx *ctx);
void *xp_ctx_stream(const xp_ctx *ctx); /* cudaStream_t */
/* Measured non-fused FP64 throughput of this device in 1e12 operations/s (one DMUL or one
 * DADD of one lane = one operation; ~50 ms).  The update kernels may not fuse the multiply
 * and the add (lpsol.h:1487-1488 rounds twice), so this -- not the DFMA peak -- is the
 * roofline bench.py reports against. */
int xp_probe_fp64_nonfused(xp_ctx *ctx, double *tops, double *ms);
/* Page-locked host memory (cudaMallocHost) so uploads/downloads of a large
 * tableau run at full PCIe rate; plain malloc'ed buffers work too, slower. */
int xp_host_alloc(xp_ctx *ctx, size_t bytes, void **out);
int xp_host_free(xp_ctx *ctx, void *p);

/* ---------------------------------------------- kernel level: large FP64 LP
 * Replaces SIX<FloatMat,Float>::solveSlackForm (lpsol.h:1007-1191) together
 * with findPivotBV (:552), findPivotNVandBVPair (:670), pivot (:1455),
 * PivotPairTab (:68) and is_feasible (:783) on an HBM-resident tableau.
 *
 * tableau m x C (C = rhs_idx+1) and tgtf 1 x C are updated in place exactly as
 * the reference leaves them; nvset/bvset (rhs_idx bytes), bv2eq (rhs_idx),
 * eq2bv (m) are the reference's Vector<bool>/Vector<INT> basis maps.
 * vc_diag/vc_rhs: the diagonal and constant column of 'vc' (rhs_idx entries,
 * the only parts is_feasible reads, lpsol.h:799); NULL means -1 / 0.
 * sol: C entries.  pivot_log (optional): up to log_cap {nv, bv, row} triples.
 */
int xp_six_slack_f64(xp_ctx *ctx, double *tableau, double *tgtf, int m, int C, uint8_t *nvset,
                     uint8_t *bvset, int32_t *bv2eq, int32_t *eq2bv, const double *vc_diag,
                     const double *vc_rhs, uint32_t max_iter, int rule, double *maxv, double *sol,
                     uint32_t *iters, int32_t *pivot_log, uint32_t log_cap);

/* Device-resident handle for the same path: upload once, run K pivots at a
 * time (resume semantics: the tabu table and iteration count persist until the
 * next upload), download state.  This is what bench.py times for `value`. */
typedef struct xp_lp_f64 xp_lp_f64;
int xp_lp_f64_create(xp_ctx *ctx, int m, int C, xp_lp_f64 **out);
void xp_lp_f64_destroy(xp_lp_f64 *lp);
int xp_lp_f64_upload(xp_lp_f64 *lp, const double *tableau, const double *tgtf,
                     const uint8_t *nvset, const uint8_t *bvset, const int32_t *bv2eq,
                     const int32_t *eq2bv, const double *vc_diag, const double *vc_rhs);
/* Build the slack form [A | I | b] of a normalised LP (leq m x (n+1), b >= 0)
 * directly on the device: SIX::slack + the identity basis of stage1
 * (lpsol.h:1405-1433, :1821-1841).  Requires C == n + m + 1.  On a sharded handle every
 * rank passes the same full `leq` and uploads only the columns that fall into its slice. */
int xp_lp_f64_upload_leq(xp_lp_f64 *lp, const double *leq, const double *tgtf, int n);
/* Fill with the synthetic dense family of SURVEY 8(d) on the device
 * (A_ij~U(0,1), b_i = 1 + U*n, c_j~U(0,1); counter-based generator). */
int xp_lp_f64_fill_synthetic(xp_lp_f64 *lp, uint64_t seed);
/* Run until a terminal status or until the total iteration count reaches
 * max_iter (SIX_TIME_OUT).  Returns the status. */
int xp_lp_f64_solve(xp_lp_f64 *lp, uint32_t max_iter, int rule);
int xp_lp_f64_download(xp_lp_f64 *lp, double *tableau, double *tgtf, uint8_t *nvset,
                       uint8_t *bvset, int32_t *bv2eq, int32_t *eq2bv, double *maxv, double *sol,
                       uint32_t *iters, int32_t *pivot_log, uint32_t log_cap);
/* Pivots per pass over the tableau (1..XP_MAX_BLOCK, 0 = automatic).  The
 * tableau in HBM is brought up to date every k pivots by one kernel that applies
 * the k pending rank-1 updates to each entry in the reference's order (same
 * roundings, hence the same bits); k = 1 is the reference's own schedule of one
 * full read+write of the tableau per pivot (lpsol.h:1481-1490).  Takes effect at
 * the next upload / solve. */
#define XP_MAX_BLOCK 32
int xp_lp_f64_set_block(xp_lp_f64 *lp, int pivots_per_flush);
int xp_ctx_set_block(xp_ctx *ctx, int pivots_per_flush); /* for xp_six_slack_f64 */
/* Pricing window of the panel kernel.  The reference enters the LOWEST-index column
 * with c_j > 0 (lpsol.h:1054-1069); while that column lies among the first `width`
 * columns, the decision chain of a pivot (entering column, ratio test, leaving row,
 * pricing) is run by one 16-CTA cluster that carries the objective row and the pivot
 * rows for that window only, and everything to its right is brought up to date once per
 * block.  Same operations per entry in the same order, hence the same bits; a pricing
 * scan that leaves the window falls back to the full-width kernels for that pivot.
 * width: 0 = automatic (on for large LPs: at most 4096 columns, and where rank 0's slice is at
 * least twice that the width follows the entering column -- 1.5 x the highest column used lately
 * + 256 -- since everything inside the window has to be in place before the next block can be
 * decided), < 0 = off, > 0 = forced.  xp_lp_f64_window() returns the width in use right now.
 * Every rank of a sharded LP must make the same call; takes effect at the next solve. */
int xp_lp_f64_set_window(xp_lp_f64 *lp, int width);
int xp_lp_f64_window(const xp_lp_f64 *lp); /* width in use (0 = off) */
int xp_ctx_set_window(xp_ctx *ctx, int width); /* for xp_six_slack_f64 / xp_six_two_stage_f64_large */
/* Per-launch timing of the tableau-update kernel (k_flush) with CUDA events on
 * the ctx stream (bench.py's roofline leg).  sweep_ms sums the launches that did
 * real work since enable; gap_ms the time between consecutive ones (the panel
 * kernels k_pcol / k_prow of the pivots in between + launch gaps). */
int xp_lp_f64_profile(xp_lp_f64 *lp, int enable);
int xp_lp_f64_profile_read(xp_lp_f64 *lp, uint64_t *n_sweeps, double *sweep_ms, double *gap_ms);
/* SMs the timed tableau passes of the last xp_lp_f64_solve shared with the deciding kernel: 0 --
 * each pass had the device to itself -- or 16 when the lookahead was on (windowed panel, on the
 * rank that runs it: a finished block is applied to the tableau beside the 16-CTA cluster that
 * decides the next one; the pass is then timed from the cluster's launch to its own end).  The environment
 * variable XP_NO_LOOKAHEAD=1 turns the lookahead off. */
int xp_lp_f64_pass_shared_sms(const xp_lp_f64 *lp);
/* Order-independent 64-bit checksum of the device tableau bits (parity at full size). */
int xp_lp_f64_checksum(xp_lp_f64 *lp, uint64_t *sum_tableau, uint64_t *sum_tgtf);
/* Column-sharded multi-GPU, one process per GPU.  Rank r of nranks
 * owns a contiguous column slice of the tableau and of the objective row; the
 * constant column, basis maps and tabu table are replicated.  Per pivot the
 * select kernel exchanges one 8-byte pricing candidate per rank (lowest index
 * wins) and the sweep kernel writes the entering column (m+1 doubles) into
 * every peer's exchange block -- NVLink peer memory mapped with CUDA IPC, no
 * host or library collective on the pivot path.  Set-up: every rank creates its
 * shard, the XP_PEER_HANDLE_BYTES handles are all-gathered out of band (any
 * channel: torch.distributed, MPI, a file), every rank attaches.  upload /
 * fill / solve / download / checksum then work as for one GPU: upload takes the
 * full arrays and keeps this rank's slice, download writes only this rank's
 * columns of `tableau` / `tgtf` (replicated state is returned whole), and the
 * per-rank checksums add up (mod 2^64) to the single-GPU checksum.  All ranks
 * must issue the same calls in the same order. */
#define XP_MAX_RANKS 8
#define XP_PEER_HANDLE_BYTES 64
int xp_lp_f64_create_sharded(xp_ctx *ctx, int m, int C, int rank, int nranks, xp_lp_f64 **out);
int xp_lp_f64_local_cols(const xp_lp_f64 *lp, int *col0, int *ncols);
int xp_lp_f64_peer_handle(xp_lp_f64 *lp, void *handle /* XP_PEER_HANDLE_BYTES */);
int xp_lp_f64_peer_attach(xp_lp_f64 *lp, const void *handles /* nranks x XP_PEER_HANDLE_BYTES */);
/* Same-process variant (several shards driven by threads of one process). */
int xp_lp_f64_peer_attach_local(xp_lp_f64 *lp, xp_lp_f64 *const *all /* nranks, rank order */);

/* ------------------------------------------ TwoStageMethod level: batched FP64
 * Replaces SIX<FloatMat,Float>::TwoStageMethod (lpsol.h:1906-1930: stage1,
 * slack, constructBasicFeasibleSolution, solveSlackForm) for a batch of
 * independent normalised LPs (x >= 0, no equalities): one warp per LP with the
 * tableau in registers (up to 32 rows x 64 variables), one CTA per LP with the
 * tableau in shared memory beyond that.  Large uniform batches are uploaded in
 * chunks while earlier chunks are being solved.
 *
 * Uniform batch: every LP has leq m x (n+1) at leq + k*m*(n+1) and tgtf 1 x (n+1)
 * at tgtf + k*(n+1).  Outputs per LP k (any may be NULL): status[k];
 * maxv[k] = tgtf[rhs] (lpsol.h:1119); slack_sol[k*ldo .. ] (ldo = n+m+1 entries);
 * tgtf_out[k*ldo ..] the final objective row (needed by minm, lpsol.h:1713-1716);
 * eq2bv[k*m ..]; iters[k] main-loop iteration count; pivots[k] all pivot() calls.
 */
int xp_six_two_stage_f64_batch(xp_ctx *ctx, int batch, int m, int n, const double *leq,
                               const double *tgtf, uint32_t max_iter, int rule, int32_t *status,
                               double *maxv, double *slack_sol, double *tgtf_out, int32_t *eq2bv,
                               uint32_t *iters, uint32_t *pivots);
/* Same with device pointers for every array (inputs already resident in HBM). */
int xp_six_two_stage_f64_batch_dev(xp_ctx *ctx, int batch, int m, int n, const double *d_leq,
                                   const double *d_tgtf, uint32_t max_iter, int rule,
                                   int32_t *d_status, double *d_maxv, double *d_slack_sol,
                                   double *d_tgtf_out, int32_t *d_eq2bv, uint32_t *d_iters,
                                   uint32_t *d_pivots);
/* Ragged batch: LP k has ms[k] rows, ns[k] variables, data at leq + leq_off[k],
 * tgtf + tgtf_off[k]; outputs use stride ldo >= max(ns+ms)+1 and ldm >= max(ms). */
int xp_six_two_stage_f64_ragged(xp_ctx *ctx, int batch, const int32_t *ms, const int32_t *ns,
                                const int64_t *leq_off, const int64_t *tgtf_off,
                                const double *leq, size_t leq_len, const double *tgtf,
                                size_t tgtf_len, uint32_t max_iter, int rule, int ldo, int ldm,
                                int32_t *status, double *maxv, double *slack_sol,
                                double *tgtf_out, int32_t *eq2bv, uint32_t *iters,
                                uint32_t *pivots);

/* One FP64 LP of any size on the HBM-resident path, same contract as one entry of
 * the batch above (slack_sol / tgtf_out: n+m+1 entries, eq2bv: m).  Phase 1
 * (constructBasicFeasibleSolution, lpsol.h:838-988: auxiliary column, forced first
 * pivot, auxiliary solve, pivoting xa out, objective restoration, column deletion)
 * runs on the device: the only bulk transfer is the upload of leq.  *status
 * receives the SIX status; the return value is 0 or a negative error.
 * A large bounded run without phase 1 uploads BEHIND the solve: the columns of the
 * pricing window go up first and the device starts deciding pivots on them (plus the
 * slack identity it generates itself) while the remaining columns of leq are still
 * crossing PCIe; those replay the decided blocks when they land.  Same operations per
 * entry in the same order (tests compare every bit), the upload is hidden behind the
 * first half of the work. */
int xp_six_two_stage_f64_large(xp_ctx *ctx, int m, int n, const double *leq, const double *tgtf,
                               uint32_t max_iter, int rule, int32_t *status, double *maxv,
                               double *slack_sol, double *tgtf_out, int32_t *eq2bv,
                               uint32_t *iters, uint32_t *pivots);

/* The same with the caller's variable constraints: vc_diag[j] = vc(j,j), vc_rhs[j] = vc(j,rhs)
 * for the n structural variables -- the only entries of `vc` the solver reads, in the
 * feasibility check of the optimal exit (lpsol.h:798-802); NULL = -1 / 0. */
int xp_six_two_stage_f64_large_vc(xp_ctx *ctx, int m, int n, const double *leq, const double *tgtf,
                                  const double *vc_diag, const double *vc_rhs, uint32_t max_iter,
                                  int rule, int32_t *status, double *maxv, double *slack_sol,
                                  double *tgtf_out, int32_t *eq2bv, uint32_t *iters, uint32_t *pivots);
/* TwoStageMethod on the explicit DUAL of a normalised primal (what SIX::minm solves,
 * lpsol.h:1661-1732 + calcDualMaxm :1585-1655), the dual built on the device: pass the
 * PRIMAL (leq mp x (np+1), tgtf np+1; x >= 0, no equalities); -A^T | c, the dual objective
 * -b and the slack form never exist on the host.  Outputs are the dual LP's (np rows, mp
 * variables): slack_sol / tgtf_out mp+np+1 entries, eq2bv np. */
int xp_six_two_stage_f64_large_dual(xp_ctx *ctx, int mp, int np, const double *leq,
                                    const double *tgtf, uint32_t max_iter, int rule,
                                    int32_t *status, double *maxv, double *slack_sol,
                                    double *tgtf_out, int32_t *eq2bv, uint32_t *iters,
                                    uint32_t *pivots);
/* What SIX::TwoStageMethod hands back through its IN OUT arguments (lpsol.h:291-301) after
 * the last xp_six_two_stage_f64_large[_vc] / xp_six_slack_f64 call on this ctx: the final
 * tableau m x C (C = n+m+1), objective row (C), nvset / bvset / bv2eq (C-1), eq2bv (m).
 * Any pointer may be NULL; the tableau is the only large transfer and is optional. */
int xp_ctx_last_lp_download(xp_ctx *ctx, double *tableau, double *tgtf, uint8_t *nvset,
                            uint8_t *bvset, int32_t *bv2eq, int32_t *eq2bv);
/* Checksums (as xp_lp_f64_checksum) of the tableau / objective row that the last
 * xp_six_two_stage_f64_large or xp_six_slack_f64 call on this ctx left on the device. */
int xp_ctx_last_lp_checksum(xp_ctx *ctx, uint64_t *sum_tableau, uint64_t *sum_tgtf);

/* ------------------------- TwoStageMethod level: batched exact (fraction-free)
 * The exact twin of SIX<RMat,Rational>::TwoStageMethod: integer tableau N with
 * one common denominator D per LP (a_ij = N_ij / D), int64 entries, 128-bit
 * products, exact division; same pivot rule / tabu table / phase 1.  Inputs are
 * integer matrices (int64).  Outputs are reduced num/den pairs (int64) equal to
 * the reference's reduced Rational whenever the reference did not overflow
 * int32 (its appro() path, rational.cpp:189-226).  status XP_ERR_OVERFLOW marks
 * an LP whose entries left int64.
 */
int xp_six_two_stage_i64_batch(xp_ctx *ctx, int batch, int m, int n, const int64_t *leq,
                               const int64_t *tgtf, uint32_t max_iter, int rule, int32_t *status,
                               int64_t *maxv_num_den /* 2 per LP */,
                               int64_t *slack_sol_num /* ldo per LP */,
                               int64_t *slack_sol_den /* ldo per LP */,
                               int64_t *tgtf_out_num, int64_t *tgtf_out_den, int32_t *eq2bv,
                               uint32_t *iters, uint32_t *pivots);
int xp_six_two_stage_i64_ragged(xp_ctx *ctx, int batch, const int32_t *ms, const int32_t *ns,
                                const int64_t *leq_off, const int64_t *tgtf_off,
                                const int64_t *leq, size_t leq_len, const int64_t *tgtf,
                                size_t tgtf_len, uint32_t max_iter, int rule, int ldo, int ldm,
                                int32_t *status, int64_t *maxv_num_den, int64_t *slack_sol_num,
                                int64_t *slack_sol_den, int64_t *tgtf_out_num,
                                int64_t *tgtf_out_den, int32_t *eq2bv, uint32_t *iters,
                                uint32_t *pivots);

/* ------------------------------------------------------------- entry level
 * Replace SIX<FloatMat,Float>::maxm / minm (lpsol.h:1992 / :1661) and
 * SIX<RMat,Rational>::maxm / minm: verify, normalize (eq -> ineq, free-variable
 * split), [explicit dual for min], TwoStageMethod on the GPU, calcFinalSolution.
 * tgtf 1 x (n+1); vc n x (n+1) or NULL (= -I | 0); eq k x (n+1) (k may be 0);
 * leq m x (n+1).  sol: n+1 entries, written on XP_SIX_SUCC.  eq2bv_out
 * (optional): the final basis of the LP that was solved, which the reference
 * only exposes through TwoStageMethod.  Capacity: m + 2k entries for maxm (the
 * normalised primal: every equality becomes two inequalities); 2n entries for
 * minm, which solves the explicit dual (lpsol.h:1585-1655: one row per
 * normalised variable, n plus at most n free-variable splits).  Never more than
 * that is written; entries past the solved LP's row count are left untouched.
 */
int xp_six_maxm_f64(xp_ctx *ctx, int m, int n, const double *tgtf, const double *vc, int k,
                    const double *eq, const double *leq, uint32_t max_iter, double *maxv,
                    double *sol, int32_t *eq2bv_out);
int xp_six_minm_f64(xp_ctx *ctx, int m, int n, const double *tgtf, const double *vc, int k,
                    const double *eq, const double *leq, uint32_t max_iter, double *minv,
                    double *sol, int32_t *eq2bv_out);
int xp_six_maxm_rat(xp_ctx *ctx, int m, int n, const xp_rat *tgtf, const xp_rat *vc, int k,
                    const xp_rat *eq, const xp_rat *leq, uint32_t max_iter, xp_rat *maxv,
                    xp_rat *sol, int32_t *eq2bv_out);
int xp_six_minm_rat(xp_ctx *ctx, int m, int n, const xp_rat *tgtf, const xp_rat *vc, int k,
                    const xp_rat *eq, const xp_rat *leq, uint32_t max_iter, xp_rat *minv,
                    xp_rat *sol, int32_t *eq2bv_out);

/* Batched entry level (uniform shape, vc = -I, no equalities): what a caller
 * with many independent LPs uses; `is_min` selects minm.  sol: batch x (n+1). */
int xp_six_solve_f64_batch(xp_ctx *ctx, int is_min, int batch, int m, int n, const double *tgtf,
                           const double *leq, uint32_t max_iter, int32_t *status, double *v,
                           double *sol);
int xp_six_solve_rat_batch(xp_ctx *ctx, int is_min, int batch, int m, int n, const xp_rat *tgtf,
                           const xp_rat *leq, uint32_t max_iter, int32_t *status, xp_rat *v,
                           xp_rat *sol);

/* MIP<Mat,T>::maxm / minm (lpsol.h:2635 / :2680): depth-first branch & bound
 * replayed on the host in the reference's order (fork_count, m_cur_best_v are
 * order dependent, lpsol.h:2474-2497) with node LP relaxations solved on the
 * GPU.  vc must be -I | 0 (MIP::verify, lpsol.h:2349-2358).  n_nodes (optional)
 * returns the number of node LPs solved. */
int xp_mip_solve_rat(xp_ctx *ctx, int is_min, int is_bin, int m, int n, const xp_rat *tgtf,
                     int k, const xp_rat *eq, const xp_rat *leq, xp_rat *v, xp_rat *sol,
                     int32_t *n_nodes);
/* The same with MIP's rational_indicator (lpsol.h:2626-2630, :2369-2391): n+1 flags, a
 * non-zero flag lets that entry of the solution stay rational (it is never branched on). */
int xp_mip_solve_rat_ri(xp_ctx *ctx, int is_min, int is_bin, int m, int n, const xp_rat *tgtf,
                        int k, const xp_rat *eq, const xp_rat *leq,
                        const uint8_t *rational_indicator, xp_rat *v, xp_rat *sol,
                        int32_t *n_nodes);
int xp_mip_solve_f64(xp_ctx *ctx, int is_min, int is_bin, int m, int n, const double *tgtf, int k,
                     const double *eq, const double *leq, double *v, double *sol,
                     int32_t *n_nodes);
/* A batch of independent MIPs of one shape (config 5: knapsack-style B&B):
 * all trees advance together, each wave of node relaxations is one batched
 * GPU call, decisions are replayed per tree in DFS order. */
int xp_mip_solve_rat_batch(xp_ctx *ctx, int is_min, int is_bin, int batch, int m, int n,
                           const xp_rat *tgtf, const xp_rat *leq, int32_t *status, xp_rat *v,
                           xp_rat *sol, int32_t *n_nodes);

/* Lineq::has_solution (linsys.cpp:830-906) for a batch of independent
 * dependence-feasibility systems of one shape (vc = -I, no equalities).
 * result[k] = 1 / 0. */
int xp_has_solution_rat_batch(xp_ctx *ctx, int batch, int m, int n, const xp_rat *leq,
                              int is_int_sol, int is_unique_sol, int32_t *result);
/* The same for the shape the real producer emits (DepPolyMgr::buildDepPoly,
 * poly.cpp:1166-1195: one query per reference pair and loop depth): system b has
 * ns[b] variables, ms[b] inequality rows at leq_pool + leq_off[b] and ks[b]
 * equality rows at eq_pool + eq_off[b] (offsets in xp_rat elements, rows of
 * ns[b]+1 entries; ks / eq_* may be NULL).  All max problems of the batch go to
 * the GPU together, then the min problems of the systems still undecided, and
 * B&B trees advance in lockstep.  result[b] = 1 / 0, or XP_ERR_REFERENCE_UB for a
 * system on which the reference itself has undefined behaviour (equalities
 * without inequalities, linsys.cpp:851-854; convertEq2Ineq, lpsol.h:1232); the
 * other systems are still answered.  A system that does not lie inside its pool
 * (negative offset, offset + rows*(n+1) > pool length) fails the call with
 * XP_ERR_BAD_ARG before anything is read. */
int xp_has_solution_rat_ragged(xp_ctx *ctx, int batch, const int32_t *ns, const int32_t *ms,
                               const int64_t *leq_off, const xp_rat *leq_pool,
                               size_t leq_pool_len /* xp_rat elements */, const int32_t *ks,
                               const int64_t *eq_off, const xp_rat *eq_pool,
                               size_t eq_pool_len /* xp_rat elements */, int is_int_sol,
                               int is_unique_sol, int32_t *result);

/* ------------------------------------------------ batches across several GPUs
 * small-LP batches, B&B trees and dependence queries are independent
 * units.  `ctxs` holds nctx contexts (normally one per device of this process; several
 * on one device also work); the batch is cut into nctx contiguous slices, every slice
 * is solved by its context on its own device, stream and host thread -- no collective --
 * and the outputs land in the caller's arrays at the slice's offset.  Arguments as in
 * the single-context calls. */
int xp_six_two_stage_f64_batch_multi(xp_ctx *const *ctxs, int nctx, int batch, int m, int n,
                                     const double *leq, const double *tgtf, uint32_t max_iter,
                                     int rule, int32_t *status, double *maxv, double *slack_sol,
                                     double *tgtf_out, int32_t *eq2bv, uint32_t *iters,
                                     uint32_t *pivots);
int xp_six_two_stage_i64_batch_multi(xp_ctx *const *ctxs, int nctx, int batch, int m, int n,
                                     const int64_t *leq, const int64_t *tgtf, uint32_t max_iter,
                                     int rule, int32_t *status, int64_t *maxv_num_den,
                                     int64_t *slack_sol_num, int64_t *slack_sol_den,
                                     int64_t *tgtf_out_num, int64_t *tgtf_out_den, int32_t *eq2bv,
                                     uint32_t *iters, uint32_t *pivots);
int xp_mip_solve_rat_batch_multi(xp_ctx *const *ctxs, int nctx, int is_min, int is_bin, int batch,
                                 int m, int n, const xp_rat *tgtf, const xp_rat *leq,
                                 int32_t *status, xp_rat *v, xp_rat *sol, int32_t *n_nodes);
int xp_has_solution_rat_ragged_multi(xp_ctx *const *ctxs, int nctx, int batch, const int32_t *ns,
                                     const int32_t *ms, const int64_t *leq_off,
                                     const xp_rat *leq_pool, size_t leq_pool_len, const int32_t *ks,
                                     const int64_t *eq_off, const xp_rat *eq_pool,
                                     size_t eq_pool_len, int is_int_sol, int is_unique_sol,
                                     int32_t *result);

#ifdef __cplusplus
}
#endif
#endif /* XPOLY_B200_H */
