"""Parity of the one-CTA-per-LP FP64 kernel (xp_six_two_stage_f64_batch /
_ragged) against the oracle's TwoStageMethod, bit for bit per LP: status,
objective, slack solution, final objective row, basis, pivot count."""
import numpy as np
import pytest

import harness as H
import xpoly_b200 as xp

pytestmark = pytest.mark.gpu


def oracle_batch(lps, max_iter=H.NO_LIMIT):
    return [H.two_stage("oracle", "f64", l, t, max_iter, want_log=True) for l, t in lps]


def check_lp(g, k, o, m, n, tag):
    assert g["status"][k] == o["status"], (tag, k, g["status"][k], o["status"])
    assert g["pivots"][k] == len(o["log"]), (tag, k, "pivots", g["pivots"][k], len(o["log"]))
    if o["status"] == H.SIX_NO_PRI:
        return
    Cc = o["cols"]
    assert Cc == n + m + 1
    assert np.array_equal(g["eq2bv"][k][:m], o["eq2bv"]), (tag, k, "eq2bv")
    assert np.array_equal(H.bits(g["tgtf"][k][:Cc]), H.bits(o["tgtf"])), (tag, k, "tgtf")
    assert np.array_equal(H.bits(g["maxv"][k:k + 1]), H.bits(o["maxv"])), (tag, k, "maxv")
    assert np.array_equal(H.bits(g["slack_sol"][k][:Cc]), H.bits(o["slack_sol"])), (tag, k, "sol")


def run_uniform(ctx, lps, max_iter=H.NO_LIMIT, tag=None):
    leq = np.stack([l for l, _ in lps])
    tg = np.stack([t for _, t in lps])
    g = ctx.two_stage_f64_batch(leq, tg, max_iter)
    m, n = leq.shape[1], leq.shape[2] - 1
    for k, o in enumerate(oracle_batch(lps, max_iter)):
        check_lp(g, k, o, m, n, tag)
    return g


def _ab(ctx, fn):
    """Run fn() with the one-warp-per-LP register kernel (default for LPs of at most 32 rows /
    64 variables) and again with the one-CTA-per-LP shared-memory kernel forced."""
    import os
    os.environ["XP_BATCH_WARP"] = "1"
    a = fn()
    os.environ["XP_BATCH_WARP"] = "0"
    try:
        b = fn()
    finally:
        os.environ.pop("XP_BATCH_WARP", None)
    return a, b


def test_c2_shape_dense_32x64(ctx):
    """Config 2 shape (tableau 32x64), SURVEY 8(d) distribution: status mix SUCC /
    OPTIMAL_IS_INFEASIBLE / UNBOUND must match per LP."""
    lps = [H.gen_dense_lp(2024 + k, 32, 31) for k in range(300)]
    g = run_uniform(ctx, lps, tag="c2")
    assert set(np.unique(g["status"])) <= {0, 1, 3}
    assert (g["status"] == 0).sum() > 50 and (g["status"] == 3).sum() > 50


@pytest.mark.parametrize("m,n", [(1, 1), (2, 2), (3, 7), (8, 7), (16, 15), (24, 23), (40, 20)])
def test_phase1_and_mixed_sign(ctx, m, n):
    """Negative constant terms / no positive cost force the auxiliary LP
    (constructBasicFeasibleSolution): forced pivot, aux solve, xa pivot-out,
    objective substitution, column deletion."""
    lps = [H.gen_mixed_lp(100 * m + k, m, n, bneg=0.3) for k in range(60)]
    lps += [H.gen_mixed_lp(7000 + 100 * m + k, m, n) for k in range(40)]
    g = run_uniform(ctx, lps, tag=("phase1", m, n))
    if m >= 8:
        assert H.SIX_NO_PRI in set(g["status"].tolist())


def test_bounded_iterations(ctx):
    lps = [H.gen_dense_lp(5 + k, 16, 15) for k in range(40)]
    for K in (0, 1, 3, 7):
        run_uniform(ctx, lps, max_iter=K, tag=("K", K))


def test_integer_data_dependence_style(ctx):
    """Small integer data (A in {-1,0,1,2}, 30% dense, b in [0,20]) resembling
    dependence polyhedra (second c2 family)."""
    lps = [H.gen_int_lp(31 + k, 12, 8, alo=-1, ahi=2, density=0.3, blo=0, bhi=20) for k in range(200)]
    run_uniform(ctx, lps, tag="dep")


def test_ragged_batch(ctx):
    r = np.random.RandomState(5)
    lps = []
    for k in range(150):
        m, n = int(r.randint(1, 20)), int(r.randint(1, 20))
        lps.append(H.gen_mixed_lp(900 + k, m, n, bneg=0.2) if k % 2 else H.gen_dense_lp(900 + k, m, n))
    g = ctx.two_stage_f64_ragged(lps)
    for k, o in enumerate(oracle_batch(lps)):
        check_lp(g, k, o, lps[k][0].shape[0], lps[k][0].shape[1] - 1, "ragged")


def test_c5_node_shape_51x102(ctx):
    """B&B node relaxation shape (knapsack row + x_j <= 1 rows, 50 variables)."""
    lps = []
    for k in range(6):
        r = np.random.RandomState(99 + k)
        n = 50
        w = r.randint(5, 41, size=n)
        p = r.randint(5, 61, size=n)
        leq = np.zeros((n + 1, n + 1))
        leq[0, :n] = w
        leq[0, n] = int(w.sum() // 3)
        for j in range(n):
            leq[1 + j, j] = 1
            leq[1 + j, n] = 1
        tg = np.zeros(n + 1)
        tg[:n] = p
        lps.append((leq, tg))
    run_uniform(ctx, lps, tag="c5")


def test_empty_batch_and_beyond_shared_memory(ctx):
    """An empty batch is a no-op; LPs whose state exceeds one SM's shared memory
    run the same kernel with the state slab in global memory (no error, same bits)."""
    out = ctx.two_stage_f64_batch(np.zeros((0, 4, 5)), np.zeros((0, 5)))
    assert out["status"].shape == (0,)
    lps = [H.gen_dense_lp(77, 130, 129), H.gen_mixed_lp(5, 130, 129, bneg=0.2)]
    run_uniform(ctx, lps, max_iter=60, tag="gmem-slab")


def test_c2_full_batch_100k(ctx):
    """BASELINE config 2 at its full size: 100 000 LPs of tableau 32 x 64 in one call.  Properties
    that do not need the oracle on every LP: (a) an LP's result does not depend on where it sits
    in the batch (reversed order gives the reversed results, bit for bit); (b) a ragged launch of
    the same LPs agrees with the uniform one; (c) a random sample of 200 LPs equals the oracle."""
    B, m, n = 100_000, 32, 31
    from xpoly_b200 import synth
    r = np.random.RandomState(2024)
    leq, tg = synth.dense_lp_batch(2024, B, m, n)  # LP k from std::mt19937_64(2024 + k)
    mix = None
    a = ctx.two_stage_f64_batch(leq, tg)
    b = ctx.two_stage_f64_batch(leq[::-1].copy(), tg[::-1].copy())
    # (d) the shared-memory CTA kernel, forced, gives the same bits on all 100 000 LPs as the
    # register-resident warp kernel that serves this shape by default
    _, cta = _ab(ctx, lambda: ctx.two_stage_f64_batch(leq, tg))
    for k in ("status", "pivots", "iters", "eq2bv"):
        assert np.array_equal(a[k], cta[k]), ("warp vs cta", k)
    for k in ("maxv", "slack_sol", "tgtf"):
        assert np.array_equal(H.bits(a[k]), H.bits(cta[k])), ("warp vs cta", k)
    for k in ("status", "pivots", "eq2bv"):
        assert np.array_equal(a[k], b[k][::-1]), k
    for k in ("maxv", "slack_sol", "tgtf"):
        assert np.array_equal(H.bits(a[k]), H.bits(b[k][::-1])), k
    assert set(np.unique(a["status"])) <= {0, 1, 3}
    mix = {k: float((a["status"] == k).mean()) for k in (0, 1, 3)}
    assert 0.40 < mix[0] < 0.60 and 0.40 < mix[3] < 0.60 and mix[1] < 0.05, mix  # SURVEY: ~49 / 50 / 1.6 %
    idx = r.choice(B, size=200, replace=False)
    rag = ctx.two_stage_f64_ragged([(leq[k], tg[k]) for k in idx[:64]])
    for j, k in enumerate(idx[:64]):
        assert rag["status"][j] == a["status"][k] and rag["pivots"][j] == a["pivots"][k]
        assert np.array_equal(H.bits(rag["maxv"][j:j + 1]), H.bits(a["maxv"][k:k + 1]))
    for k in idx:
        o = H.two_stage("oracle", "f64", leq[k], tg[k], want_log=True)
        check_lp(a, int(k), o, m, n, "c2-full")


@pytest.mark.parametrize("m,n,bneg", [(32, 31, 0.0), (32, 31, 0.3), (24, 23, 0.3), (16, 40, 0.2),
                                      (8, 55, 0.3), (8, 7, 0.3), (16, 15, 0.0), (5, 3, 0.5)])
def test_warp_kernel_equals_cta_kernel(ctx, m, n, bneg):
    """The register-resident warp kernel and the shared-memory CTA kernel are two schedules of
    the same arithmetic: every output of a 4096-LP batch must agree bit for bit, for every
    instantiation (rows 8/16/24/32, one or two column slots), with and without phase 1."""
    B = 4096
    r = np.random.RandomState(1000 * m + n)
    leq = r.uniform(-1 if bneg else 0, 1, size=(B, m, n + 1))
    leq[:, :, n] = np.where(r.uniform(size=(B, m)) < bneg, -1.0, 1.0) * (1.0 + r.uniform(size=(B, m)) * n)
    tg = r.uniform(-0.2 if bneg else 0, 1, size=(B, n + 1))
    tg[:, n] = 0.0
    for K in (xp.NO_ITER_LIMIT, 5):
        a, b = _ab(ctx, lambda: ctx.two_stage_f64_batch(leq, tg, K))
        assert np.array_equal(a["status"], b["status"])
        for k in ("pivots", "iters"):
            assert np.array_equal(a[k], b[k]), k
        ok = a["status"] != H.SIX_NO_PRI
        assert np.array_equal(a["eq2bv"][ok], b["eq2bv"][ok])
        for k in ("maxv", "slack_sol", "tgtf"):
            assert np.array_equal(H.bits(a[k][ok]), H.bits(b[k][ok])), (k, K)


def test_warp_kernel_ragged_equals_cta_kernel(ctx):
    r = np.random.RandomState(11)
    lps = []
    for k in range(600):
        m, n = int(r.randint(1, 33)), int(r.randint(1, 32))
        if n + 1 + m > 64:
            n = 63 - m
        lps.append(H.gen_mixed_lp(4000 + k, m, n, bneg=0.2) if k % 2 else H.gen_dense_lp(4000 + k, m, n))
    a, b = _ab(ctx, lambda: ctx.two_stage_f64_ragged(lps))
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["pivots"], b["pivots"])
    ok = a["status"] != H.SIX_NO_PRI
    assert np.array_equal(a["eq2bv"][ok], b["eq2bv"][ok])
    for k in ("maxv", "slack_sol", "tgtf"):
        assert np.array_equal(H.bits(a[k][ok]), H.bits(b[k][ok])), k


def lower_bound_lp(seed, m, n, nneg, integer=False):
    """Feasible LPs that need phase 1 and pass it (dense family + lower bounds -x_j <= -l_j)."""
    r = np.random.RandomState(seed)
    leq = np.zeros((m + nneg, n + 1))
    if integer:
        leq[:m, :n] = r.randint(0, 4, size=(m, n)) * (r.uniform(size=(m, n)) < 0.5)
        leq[:m, n] = r.randint(8, 30, size=m)
    else:
        leq[:m, :n] = r.uniform(0, 1, size=(m, n))
        leq[:m, n] = 1 + r.uniform(0, 1, size=m) * n
    for t in range(nneg):
        leq[m + t, t] = -1.0
        leq[m + t, n] = -1.0 if integer else -0.01 * (t + 1)
    tg = np.zeros(n + 1)
    tg[:n] = r.randint(1, 6, size=n) if integer else r.uniform(0, 1, size=n)
    return leq, tg


@pytest.mark.parametrize("m,n,nneg", [(5, 4, 1), (20, 15, 2), (28, 30, 3)])
def test_phase1_succeeds_objective_restored(ctx, m, n, nneg):
    """Phase 1 that succeeds -- aux optimum 0, xa pivoted out, objective restored by
    substitution, column xa dropped, main solve -- on the register-resident warp kernel and,
    forced, on the CTA kernel, against the oracle."""
    lps = [lower_bound_lp(s, m, n, nneg) for s in range(120)]
    import os
    for force in ("1", "0"):
        os.environ["XP_BATCH_WARP"] = force
        try:
            g = run_uniform(ctx, lps, tag=("lb", m, n, force))
            run_uniform(ctx, lps[:40], max_iter=6, tag=("lbK", m, n, force))
        finally:
            os.environ.pop("XP_BATCH_WARP", None)
        assert H.SIX_NO_PRI not in set(g["status"].tolist())


def test_batches_split_across_contexts(ctx):
    """xp_*_batch_multi: the batch is cut into contiguous slices, one context (device, stream, host
    thread) per slice, no collective.  Contexts on every visible GPU (three on cuda:0 where there is
    only one): bit-identical to the single-context call."""
    import ctypes as C
    import torch
    ndev = torch.cuda.device_count()
    devs = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    ctxs = [xp.Context(d) for d in devs]
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    lib = xp.lib()
    P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    B, m, n = 1001, 12, 9
    lps = [H.gen_dense_lp(9100 + k, m, n) for k in range(B)]
    leq = np.ascontiguousarray(np.stack([l for l, _ in lps]))
    tg = np.ascontiguousarray(np.stack([t for _, t in lps]))
    one = ctx.two_stage_f64_batch(leq, tg)
    st, mv = np.zeros(B, dtype=np.int32), np.zeros(B)
    sol, tgo = np.zeros((B, n + m + 1)), np.zeros((B, n + m + 1))
    e2b, it, pv = np.zeros((B, m), dtype=np.int32), np.zeros(B, dtype=np.uint32), np.zeros(B, dtype=np.uint32)
    rc = lib.xp_six_two_stage_f64_batch_multi(arr, len(ctxs), B, m, n, P(leq), P(tg), C.c_uint32(xp.NO_ITER_LIMIT), 0,
                                              P(st), P(mv), P(sol), P(tgo), P(e2b), P(it), P(pv))
    assert rc == 0
    assert np.array_equal(st, one["status"]) and np.array_equal(H.bits(mv), H.bits(one["maxv"]))
    assert np.array_equal(H.bits(sol), H.bits(one["slack_sol"])) and np.array_equal(e2b, one["eq2bv"])
    assert np.array_equal(pv, one["pivots"])
    # exact LPs
    ilps = [H.gen_int_lp(9200 + k, 7, 4, alo=-1, ahi=3, density=0.7, blo=0, bhi=15) for k in range(B)]
    il = np.ascontiguousarray(np.stack([l for l, _ in ilps]).astype(np.int64))
    itg = np.ascontiguousarray(np.stack([t for _, t in ilps]).astype(np.int64))
    e1 = ctx.two_stage_i64_batch(il, itg)
    st2, mv2 = np.zeros(B, dtype=np.int32), np.zeros((B, 2), dtype=np.int64)
    rc = lib.xp_six_two_stage_i64_batch_multi(arr, len(ctxs), B, 7, 4, P(il), P(itg), C.c_uint32(xp.NO_ITER_LIMIT), 0,
                                              P(st2), P(mv2), None, None, None, None, None, None, None)
    assert rc == 0 and np.array_equal(st2, e1["status"]) and np.array_equal(mv2, e1["maxv"])
    # B&B trees and dependence queries
    T = 37
    mip1 = ctx.mip_solve_rat_batch(0, 0, il[:T], itg[:T])
    rl, rt = xp._rat(il[:T]), xp._rat(itg[:T])
    ms_, mv_ = np.zeros(T, dtype=np.int32), np.zeros((T, 2), dtype=np.int32)
    msol, mn = np.zeros((T, 5, 2), dtype=np.int32), np.zeros(T, dtype=np.int32)
    rc = lib.xp_mip_solve_rat_batch_multi(arr, len(ctxs), 0, 0, T, 7, 4, P(rt), P(rl), P(ms_), P(mv_), P(msol), P(mn))
    assert rc == 0 and np.array_equal(ms_, mip1["status"]) and np.array_equal(mv_, mip1["v"])
    assert np.array_equal(mn, mip1["nodes"])
    systems = [(il[k], None) for k in range(200)]
    h1 = ctx.has_solution_ragged(systems)
    ns = np.full(200, 4, dtype=np.int32)
    msq = np.full(200, 7, dtype=np.int32)
    off = (np.arange(200, dtype=np.int64) * 35)
    pool = np.ascontiguousarray(xp._rat(il[:200]).reshape(-1, 2))
    res = np.zeros(200, dtype=np.int32)
    rc = lib.xp_has_solution_rat_ragged_multi(arr, len(ctxs), 200, P(ns), P(msq), P(off), P(pool), C.c_size_t(len(pool)),
                                              None, None, None, C.c_size_t(0), 1, 1, P(res))
    assert rc == 0 and np.array_equal(res, h1)
    for c in ctxs:
        c.close()
