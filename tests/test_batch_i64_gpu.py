"""Parity of the exact fraction-free kernel (xp_six_two_stage_i64_batch) against
the oracle's SIX<RMat,Rational>::TwoStageMethod: identical status, pivot count
and basis; objective, solution and final objective row equal as reduced
num/den -- on every LP where the reference stayed exact (appro count 0)."""
from math import gcd

import numpy as np
import pytest

import harness as H
import xpoly_b200 as xp

pytestmark = pytest.mark.gpu


def canon(pair):
    a, b = int(pair[0]), int(pair[1])
    if a == 0:
        return (0, 1)
    g = gcd(a, b)
    a, b = a // g, b // g
    return (-a, -b) if b < 0 else (a, b)


def oracle_exact(leq, tg, max_iter=H.NO_LIMIT):
    a0 = H.appro_count("oracle")
    o = H.two_stage("oracle", "rat", H.to_rat(leq), H.to_rat(tg), max_iter, want_log=True)
    o["exact"] = H.appro_count("oracle") == a0
    return o


def check(g, k, o, m, n, tag):
    if not o["exact"]:
        return False  # the reference itself went through appro(): outside the parity set
    assert g["status"][k] == o["status"], (tag, k, g["status"][k], o["status"])
    assert g["pivots"][k] == len(o["log"]), (tag, k, "pivots")
    if o["status"] == H.SIX_NO_PRI:
        return True
    Cc = o["cols"]
    assert np.array_equal(g["eq2bv"][k][:m], o["eq2bv"]), (tag, k, "eq2bv")
    assert canon(g["maxv"][k]) == canon(o["maxv"]), (tag, k, "maxv", g["maxv"][k], o["maxv"])
    for j in range(Cc):
        assert (int(g["tgtf_num"][k][j]), int(g["tgtf_den"][k][j])) == canon(o["tgtf"][j]), \
            (tag, k, "tgtf", j)
    if o["status"] in (0, 3):
        for j in range(Cc):
            assert (int(g["sol_num"][k][j]), int(g["sol_den"][k][j])) == canon(o["slack_sol"][j]), \
                (tag, k, "sol", j)
    return True


def run(ctx, lps, max_iter=H.NO_LIMIT, tag=None):
    leq = np.stack([l for l, _ in lps])
    tg = np.stack([t for _, t in lps])
    g = ctx.two_stage_i64_batch(leq, tg, max_iter)
    m, n = leq.shape[1], leq.shape[2] - 1
    checked = sum(check(g, k, oracle_exact(l, t, max_iter), m, n, tag) for k, (l, t) in enumerate(lps))
    return g, checked


def test_c4_family_24x48(ctx):
    """Config 4: leq 24x24 (23 vars), integer A in [0,3] at 30% density,
    b in [0,20], c in [1,5]."""
    lps = [H.gen_int_lp(777 + k, 24, 23) for k in range(200)]
    g, checked = run(ctx, lps, tag="c4")
    assert checked >= 150
    assert set(np.unique(g["status"])) <= {0, 1, 3}


def test_mixed_sign_8x7_unbounded_by_exhaustion(ctx):
    lps = [H.gen_int_lp(k, 8, 7, alo=-3, ahi=3, density=0.5) for k in range(300)]
    g, checked = run(ctx, lps, tag="8x7")
    assert checked >= 250 and (g["status"] == 1).sum() > 50


def test_phase1_negative_rhs(ctx):
    lps = [H.gen_int_lp(k, 8, 7, alo=-2, ahi=3, density=0.6, blo=-5, bhi=15) for k in range(300)]
    g, checked = run(ctx, lps, tag="phase1")
    assert checked >= 250
    assert {0, 2} <= set(np.unique(g["status"]).tolist())


def test_bounded_iterations(ctx):
    lps = [H.gen_int_lp(50 + k, 12, 11) for k in range(60)]
    for K in (0, 1, 4):
        run(ctx, lps, max_iter=K, tag=("K", K))


def test_dependence_style_and_ragged(ctx):
    r = np.random.RandomState(11)
    lps = []
    for k in range(120):
        m, n = int(r.randint(2, 14)), int(r.randint(1, 9))
        lps.append(H.gen_int_lp(400 + k, m, n, alo=-1, ahi=2, density=0.4, blo=-3, bhi=20))
    g = ctx.two_stage_i64_ragged(lps)
    n_ok = 0
    for k, (l, t) in enumerate(lps):
        n_ok += check(g, k, oracle_exact(l, t), l.shape[0], l.shape[1] - 1, "ragged")
    assert n_ok >= 100


def test_overflow_is_detected_not_wrapped(ctx):
    """Large coefficients leave int64 after a few pivots: the kernel must say so."""
    r = np.random.RandomState(3)
    leq = np.zeros((4, 12, 12), dtype=np.int64)
    leq[:, :, :11] = r.randint(1 << 28, 1 << 30, size=(4, 12, 11))
    leq[:, :, 11] = r.randint(1 << 28, 1 << 30, size=(4, 12))
    tg = np.zeros((4, 12), dtype=np.int64)
    tg[:, :11] = r.randint(1 << 20, 1 << 22, size=(4, 11))
    g = ctx.two_stage_i64_batch(leq, tg)
    assert (g["status"] == xp.ERR_OVERFLOW).any()
    assert set(np.unique(g["status"]).tolist()) <= {xp.ERR_OVERFLOW, 0, 1, 3}


def test_c4_full_batch_10k(ctx):
    """BASELINE config 4 at its full size: 10 000 exact LPs of tableau 24 x 48.  Order
    independence over the whole batch, and a sample of 150 LPs against the oracle's Rational
    solver (LPs where the reference fell into appro() are outside the parity set)."""
    B, m, n = 10_000, 24, 23
    r = np.random.RandomState(777)
    A = r.randint(0, 4, size=(B, m, n)) * (r.uniform(size=(B, m, n)) < 0.3)
    leq = np.zeros((B, m, n + 1), dtype=np.int64)
    leq[:, :, :n] = A
    leq[:, :, n] = r.randint(0, 21, size=(B, m))
    tg = np.zeros((B, n + 1), dtype=np.int64)
    tg[:, :n] = r.randint(1, 6, size=(B, n))
    a = ctx.two_stage_i64_batch(leq, tg)
    b = ctx.two_stage_i64_batch(leq[::-1].copy(), tg[::-1].copy())
    for k in ("status", "pivots", "eq2bv", "maxv", "sol_num", "sol_den", "tgtf_num", "tgtf_den"):
        assert np.array_equal(a[k], b[k][::-1]), k
    assert (a["status"] == xp.ERR_OVERFLOW).sum() == 0
    checked = 0
    for k in r.choice(B, size=150, replace=False):
        checked += check(a, int(k), oracle_exact(leq[k].astype(float), tg[k].astype(float)), m, n, "c4-full")
    assert checked > 100


@pytest.mark.parametrize("m,n,neg", [(24, 23, False), (24, 23, True), (32, 31, False), (16, 40, True),
                                     (8, 55, True), (7, 4, True), (16, 15, True), (3, 2, True)])
def test_warp_kernel_equals_cta_kernel_exact(ctx, m, n, neg):
    """The register-resident warp kernel and the shared-memory CTA kernel of the exact path are
    two schedules of the same integer arithmetic: every output of a 4096-LP batch agrees, for
    every instantiation, with and without phase 1, overflow flags included."""
    import os
    B = 4096
    r = np.random.RandomState(77 * m + n)
    A = r.randint(-2 if neg else 0, 4, size=(B, m, n)) * (r.uniform(size=(B, m, n)) < 0.35)
    leq = np.zeros((B, m, n + 1), dtype=np.int64)
    leq[:, :, :n] = A
    leq[:, :, n] = r.randint(-5 if neg else 0, 21, size=(B, m))
    tg = np.zeros((B, n + 1), dtype=np.int64)
    tg[:, :n] = r.randint(-1 if neg else 1, 6, size=(B, n))
    for K in (xp.NO_ITER_LIMIT, 6):
        os.environ["XP_BATCH_WARP"] = "2"  # force the warp kernel for every shape that fits
        a = ctx.two_stage_i64_batch(leq, tg, K)
        os.environ["XP_BATCH_WARP"] = "0"
        try:
            b = ctx.two_stage_i64_batch(leq, tg, K)
        finally:
            os.environ.pop("XP_BATCH_WARP", None)
        assert np.array_equal(a["status"], b["status"]), (K, np.flatnonzero(a["status"] != b["status"])[:5])
        assert np.array_equal(a["pivots"], b["pivots"])
        assert np.array_equal(a["iters"], b["iters"])
        ok = (a["status"] >= 0) & (a["status"] != H.SIX_NO_PRI)
        for k in ("eq2bv", "maxv", "sol_num", "sol_den", "tgtf_num", "tgtf_den"):
            assert np.array_equal(a[k][ok], b[k][ok]), (k, K)


@pytest.mark.parametrize("m,n,nneg", [(5, 4, 1), (6, 8, 2), (20, 15, 2)])
def test_phase1_succeeds_exact(ctx, m, n, nneg):
    """Exact path, phase 1 that succeeds (objective restored, column xa dropped, main solve): the
    warp kernel (forced for every shape) and the CTA kernel against the Rational oracle."""
    import os
    from test_batch_f64_gpu import lower_bound_lp
    lps = [tuple(a.astype(np.int64) for a in lower_bound_lp(s, m, n, nneg, integer=True)) for s in range(100)]
    for force in ("2", "0"):
        os.environ["XP_BATCH_WARP"] = force
        try:
            g, checked = run(ctx, lps, tag=("lb-exact", m, n, force))
        finally:
            os.environ.pop("XP_BATCH_WARP", None)
        assert checked >= 50
        assert (g["status"] != H.SIX_NO_PRI).sum() >= 50
