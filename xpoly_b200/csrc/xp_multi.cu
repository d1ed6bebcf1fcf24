// Batches split across several GPUs of one process (small-LP batches, B&B trees and
// dependence queries are independent units -- no collective, every context solves a contiguous
// slice on its own device and stream, driven by its own host thread).  A context per device;
// several contexts on one device work too (tests on a single-GPU box).
#include "xp_common.cuh"

#include <thread>
#include <vector>

namespace {

template <class F>
int run_slices(xp_ctx *const *ctxs, int nctx, int batch, F f)
{
    if (!ctxs || nctx < 1 || batch < 0) return XP_ERR_BAD_ARG;
    for (int c = 0; c < nctx; c++)
        if (!ctxs[c]) return XP_ERR_BAD_ARG;
    std::vector<int> rc(nctx, 0);
    std::vector<std::thread> th;
    for (int c = 0; c < nctx; c++) {
        const int lo = (int)((long long)batch * c / nctx), hi = (int)((long long)batch * (c + 1) / nctx);
        if (hi <= lo) continue;
        th.emplace_back([&, c, lo, hi]() { rc[c] = f(ctxs[c], lo, hi - lo); });
    }
    for (auto &t : th) t.join();
    for (int c = 0; c < nctx; c++)
        if (rc[c]) return rc[c];
    return 0;
}

template <class T>
T *at(T *p, size_t off)
{
    return p ? p + off : nullptr;
}

} // namespace

extern "C" int xp_six_two_stage_f64_batch_multi(xp_ctx *const *ctxs, int nctx, int batch, int m, int n,
                                                const double *leq, const double *tgtf, uint32_t max_iter, int rule,
                                                int32_t *status, double *maxv, double *slack_sol, double *tgtf_out,
                                                int32_t *eq2bv, uint32_t *iters, uint32_t *pivots)
{
    if (m < 1 || n < 1 || !leq || !tgtf) return XP_ERR_BAD_ARG;
    const size_t ldo = (size_t)n + m + 1;
    return run_slices(ctxs, nctx, batch, [&](xp_ctx *c, int lo, int cnt) {
        return xp_six_two_stage_f64_batch(c, cnt, m, n, leq + (size_t)lo * m * (n + 1), tgtf + (size_t)lo * (n + 1),
                                          max_iter, rule, at(status, lo), at(maxv, lo), at(slack_sol, lo * ldo),
                                          at(tgtf_out, lo * ldo), at(eq2bv, (size_t)lo * m), at(iters, lo), at(pivots, lo));
    });
}

extern "C" int xp_six_two_stage_i64_batch_multi(xp_ctx *const *ctxs, int nctx, int batch, int m, int n,
                                                const int64_t *leq, const int64_t *tgtf, uint32_t max_iter, int rule,
                                                int32_t *status, int64_t *maxv_num_den, int64_t *slack_sol_num,
                                                int64_t *slack_sol_den, int64_t *tgtf_out_num, int64_t *tgtf_out_den,
                                                int32_t *eq2bv, uint32_t *iters, uint32_t *pivots)
{
    if (m < 1 || n < 1 || !leq || !tgtf) return XP_ERR_BAD_ARG;
    const size_t ldo = (size_t)n + m + 1;
    return run_slices(ctxs, nctx, batch, [&](xp_ctx *c, int lo, int cnt) {
        return xp_six_two_stage_i64_batch(c, cnt, m, n, leq + (size_t)lo * m * (n + 1), tgtf + (size_t)lo * (n + 1),
                                          max_iter, rule, at(status, lo), at(maxv_num_den, (size_t)2 * lo),
                                          at(slack_sol_num, lo * ldo), at(slack_sol_den, lo * ldo),
                                          at(tgtf_out_num, lo * ldo), at(tgtf_out_den, lo * ldo),
                                          at(eq2bv, (size_t)lo * m), at(iters, lo), at(pivots, lo));
    });
}

extern "C" int xp_mip_solve_rat_batch_multi(xp_ctx *const *ctxs, int nctx, int is_min, int is_bin, int batch, int m,
                                            int n, const xp_rat *tgtf, const xp_rat *leq, int32_t *status, xp_rat *v,
                                            xp_rat *sol, int32_t *n_nodes)
{
    if (m < 1 || n < 1 || !leq || !tgtf) return XP_ERR_BAD_ARG;
    return run_slices(ctxs, nctx, batch, [&](xp_ctx *c, int lo, int cnt) {
        return xp_mip_solve_rat_batch(c, is_min, is_bin, cnt, m, n, tgtf + (size_t)lo * (n + 1),
                                      leq + (size_t)lo * m * (n + 1), at(status, lo), at(v, lo),
                                      at(sol, (size_t)lo * (n + 1)), at(n_nodes, lo));
    });
}

extern "C" int xp_has_solution_rat_ragged_multi(xp_ctx *const *ctxs, int nctx, int batch, const int32_t *ns,
                                                const int32_t *ms, const int64_t *leq_off, const xp_rat *leq_pool,
                                                size_t leq_pool_len, const int32_t *ks, const int64_t *eq_off,
                                                const xp_rat *eq_pool, size_t eq_pool_len, int is_int_sol,
                                                int is_unique_sol, int32_t *result)
{
    if (!ns || !ms || !result) return XP_ERR_BAD_ARG;
    // offsets are absolute into the pools, so every slice passes the whole pools
    return run_slices(ctxs, nctx, batch, [&](xp_ctx *c, int lo, int cnt) {
        return xp_has_solution_rat_ragged(c, cnt, ns + lo, ms + lo, at(leq_off, lo), leq_pool, leq_pool_len,
                                          at(ks, lo), at(eq_off, lo), eq_pool, eq_pool_len, is_int_sol, is_unique_sol,
                                          result + lo);
    });
}
