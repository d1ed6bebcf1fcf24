"""Parity of the HBM-resident FP64 path (xp_six_slack_f64 / xp_lp_f64_*) against
the oracle's solveSlackForm, bit for bit: status, iteration count, pivot
sequence, basis maps, whole tableau, objective row, solution."""
import os

import numpy as np
import pytest

import harness as H
import xpoly_b200 as xp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def assert_same_state(g, o, tag):
    assert g["status"] == o["status"], (tag, g["status"], o["status"])
    assert g["iters"] == o["iters"], (tag, g["iters"], o["iters"])
    assert np.array_equal(g["log"], o["log"][: len(g["log"])]), (tag, "pivot sequence")
    for k in ("eq2bv", "bv2eq", "nvset", "bvset"):
        assert np.array_equal(g[k], o[k]), (tag, k)
    for k in ("tab", "tgtf", "maxv", "sol"):
        assert np.array_equal(H.bits(g[k]), H.bits(o[k])), (tag, k)


BLOCKS = (0, 1, 3, 32)  # pivots per tableau pass: automatic, the reference's schedule, odd, maximum


def run_both(ctx, leq, tgtf, max_iter=H.NO_LIMIT, tag=None, blocks=BLOCKS, windows=(0,)):
    sf = xp.slack_form(leq, tgtf)
    o = H.slack_solve_oracle("f64", *sf, max_iter=max_iter)
    try:
        for w in windows:
            ctx.set_window(w)
            for k in blocks:
                ctx.set_block(k)
                g = ctx.six_slack_f64(*sf, max_iter=max_iter, log_cap=1 << 16)
                assert_same_state(g, o, (tag, "block", k, "window", w))
    finally:
        ctx.set_block(0)
        ctx.set_window(0)
    return g


# pricing windows: off, one column (every scan leaves it), a few columns (scans leave it now and
# then: the full-width kernels take those pivots), the whole tableau
WINDOWS = (-1, 1, 6, 1 << 20)


@pytest.mark.parametrize("m,n", [(6, 5), (16, 15), (33, 20), (7, 40), (40, 64), (130, 129)])
def test_windowed_panel_dense(ctx, m, n):
    """k_wpanel + k_prow_bulk (one 16-CTA cluster decides inside the pricing window, the columns
    outside follow once per block) leave the oracle's bits: whole tableau, objective row, basis,
    pivot sequence -- for windows that never, sometimes and always contain the entering column."""
    for seed in range(4):
        leq, tg = H.gen_dense_lp(8100 + seed, m, n)
        run_both(ctx, leq, tg, tag=("wdense", m, n, seed), blocks=(0, 1, 5, 32), windows=WINDOWS)


@pytest.mark.parametrize("m,n", [(6, 5), (10, 9), (16, 15), (12, 30)])
def test_windowed_panel_mixed_sign(ctx, m, n):
    """Ratio test failing inside the windowed kernel (-> disableNV on the slow path), its pass 2
    (negative pivots) and tabu exhaustion, with the window forced on."""
    seen = set()
    for seed in range(8):
        leq, tg = H.gen_mixed_lp(100 + seed, m, n)
        leq[:, n] = np.abs(leq[:, n])
        seen.add(run_both(ctx, leq, tg, tag=("wmixed", m, n, seed), blocks=(0, 32), windows=(3, 1 << 20))["status"])
    assert 1 in seen


def test_windowed_panel_bounded_and_resume(ctx):
    """max_iter stops inside a windowed block; resuming continues it (factors reloaded)."""
    leq, tg = H.gen_dense_lp(8200, 96, 95)
    sf = xp.slack_form(leq, tg)
    for w in (8, 1 << 20):
        lp = ctx.large_lp(*sf[0].shape)
        lp.set_window(w)
        assert lp.window == min(w, sf[0].shape[1])
        lp.set_block(32)
        lp.upload(*sf)
        for K in (3, 10, 45, 46, 120):
            st = lp.solve(K)
            g = lp.download(log_cap=1 << 16)
            g["status"] = st
            assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=K), ("wresume", w, K))
        lp.close()


@pytest.mark.parametrize("m,n,window", [(300, 1501, 512), (257, 1300, 256), (200, 2101, 768), (130, 701, 512)])  # m + n odd: even row stride
def test_lookahead_flush(ctx, m, n, window, monkeypatch):
    """Lookahead: the live pass of k_flush_w closes a block on the window tiles only, the tiles
    beyond take it one step later beside the next block's k_wpanel (programmatic dependent
    launch).  Whole state against the oracle -- runs to termination, runs the window cannot decide
    alone (mixed signs: the slow path and the full-width panel work on the caught-up tableau),
    bounded runs resumed in the middle of a block -- and against the same run with the lookahead off."""
    for seed in range(2):
        leq, tg = H.gen_dense_lp(8600 + seed, m, n)
        run_both(ctx, leq, tg, max_iter=600, tag=("look", m, n, seed), blocks=(0, 12, 32), windows=(window,))
    leq, tg = H.gen_mixed_lp(8650, m, n)
    leq[:, n] = np.abs(leq[:, n])
    tg[:n] = np.abs(tg[:n])
    run_both(ctx, leq, tg, max_iter=600, tag=("look-mixed", m, n), blocks=(0, 32), windows=(window,))
    for seed in (8651, 8652):  # mixed-sign objective as well: ratio tests fail inside the window run (-> disableNV)
        leq, tg = H.gen_mixed_lp(seed, m, n)
        leq[:, n] = np.abs(leq[:, n])
        run_both(ctx, leq, tg, max_iter=400, tag=("look-mixed2", m, n, seed), blocks=(0, 32), windows=(window,))
    leq, tg = H.gen_dense_lp(8660, m, n)
    sf = xp.slack_form(leq, tg)
    sums = {}
    for off in (False, True):
        if off:
            monkeypatch.setenv("XP_NO_LOOKAHEAD", "1")
        lp = ctx.large_lp(*sf[0].shape)
        lp.set_window(window)
        lp.set_block(32)
        lp.upload(*sf)
        for K in (3, 40, 45, 100, 101, 170):
            st = lp.solve(K)
            g = lp.download(log_cap=1 << 16)
            g["status"] = st
            assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=K), ("look-resume", off, K))
        sums[off] = lp.checksum()
        # bounded calls back to back: each leaves its last block owed, the next call applies it
        # beside its first k_wpanel; only the download at the end drains
        lp.upload(*sf)
        for K in (40, 80, 81, 120):
            st = lp.solve(K)
        g = lp.download(log_cap=1 << 16)
        g["status"] = st
        assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=120), ("look-chain", off))
        lp.close()
    assert sums[False] == sums[True]


def test_c3_window_follows_entering_column(ctx, monkeypatch):
    """Automatic window at full size: it starts at the configured 4096 columns, follows the entering
    column down (1.5 x the highest column seen lately + 256) and up again; the run is the same,
    bit for bit, as with the width fixed (checksum of all 134 M doubles after every call)."""
    m, n = 8192, 8191
    sums, widths = {}, {}
    for fixed in (False, True):
        if fixed:
            monkeypatch.setenv("XP_WINDOW_FIXED", "1")
        lp = ctx.large_lp(m, n + m + 1)
        lp.fill_synthetic(2024)
        sums[fixed], widths[fixed] = [], []
        for K in range(200, 2001, 200):
            assert lp.solve(K) == xp.SIX_TIME_OUT
            widths[fixed].append(lp.window)
            sums[fixed].append(lp.checksum())
        lp.close()
    assert sums[False] == sums[True]
    assert set(widths[True]) == {4096}
    assert min(widths[False]) <= 2048 and max(widths[False]) <= 4096, widths[False]


def test_lookahead_serialised_launches():
    """What a profiler that serialises kernels makes of the lookahead: the pass cannot start
    beside the cluster, the cluster gives up waiting for it (20 ms) without having touched anything,
    the pass runs behind it and the next k_wpanel decides the block.  Slower, same bits.  (Own
    process: the switch is read once.)"""
    import subprocess
    import sys
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import harness as H, xpoly_b200 as xp\n"
        "ctx = xp.Context(0)\n"
        "leq, tg = H.gen_dense_lp(8600, 257, 1300)\n"
        "leq2, tg2 = H.gen_dense_lp(8660, 257, 1300)\n"
        "for (l, t, K) in ((leq, tg, 600), (leq2, tg2, 100)):\n"
        "    sf = xp.slack_form(l, t)\n"
        "    o = H.slack_solve_oracle('f64', *sf, max_iter=K)\n"
        "    lp = ctx.large_lp(*sf[0].shape); lp.set_window(256); lp.set_block(16); lp.upload(*sf)\n"
        "    st = lp.solve(K); g = lp.download(log_cap=1 << 16)\n"
        "    assert st == o['status'] and g['iters'] == o['iters'], (st, o['status'], g['iters'], o['iters'])\n"
        "    for k in ('tab', 'tgtf', 'sol'):\n"
        "        assert np.array_equal(H.bits(g[k]), H.bits(o[k])), k\n"
        "    assert np.array_equal(g['eq2bv'], o['eq2bv'])\n"
        "print('serial-ok')\n") % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, XP_LOOKAHEAD_SERIAL="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "serial-ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("m,n", [(2, 2), (6, 5), (8, 7), (10, 9), (16, 15), (33, 20), (7, 40)])
def test_small_dense_to_termination(ctx, m, n):
    seen = set()
    for seed in range(12):
        leq, tg = H.gen_dense_lp(5000 + seed, m, n)
        seen.add(run_both(ctx, leq, tg, tag=("dense", m, n, seed))["status"])
    assert seen <= {0, 1, 3}


@pytest.mark.parametrize("m,n", [(6, 5), (10, 9), (16, 15), (12, 30)])
def test_mixed_sign_exhausts_tabu_table(ctx, m, n):
    """Mixed-sign data drives the retry (disableNV), pass-2 ratio test and
    findPivotNVandBVPair fallback paths; unbounded LPs exit by tabu exhaustion."""
    seen = set()
    for seed in range(10):
        leq, tg = H.gen_mixed_lp(seed, m, n)
        leq[:, n] = np.abs(leq[:, n])
        seen.add(run_both(ctx, leq, tg, tag=("mixed", m, n, seed))["status"])
    assert 1 in seen


@pytest.mark.parametrize("K", [0, 1, 2, 5, 17])
def test_bounded_iterations_state(ctx, K):
    leq, tg = H.gen_dense_lp(777, 24, 23)
    g = run_both(ctx, leq, tg, max_iter=K, tag=("K", K))
    assert g["status"] in (0, 3, 4)


def test_odd_column_count_scalar_path(ctx):
    for seed in range(6):
        leq, tg = H.gen_dense_lp(900 + seed, 9, 8)  # C = 8 + 9 + 1 = 18 even
        run_both(ctx, leq, tg, tag=("even", seed))
        leq, tg = H.gen_dense_lp(900 + seed, 9, 9)  # C = 19 odd
        run_both(ctx, leq, tg, tag=("odd", seed))


def test_c1_256x512_golden(ctx):
    """256x255 seeded instance, 14 pivots, SIX_SUCC."""
    leq, tg = H.gen_dense_lp(12345, 256, 255)
    g = run_both(ctx, leq, tg, tag="c1")
    assert g["status"] == 0 and g["iters"] == 14
    assert [tuple(r) for r in g["log"][:5]] == [(0, 493, 238), (1, 408, 153), (2, 302, 47),
                                                (5, 1, 153), (29, 2, 47)]


def test_medium_wandering_regime(ctx):
    """512 x 1024 tableau, 120 pivots: past pivot ~26 the tabu rule leaves the
    feasible region; state must still match bit for bit."""
    leq, tg = H.gen_dense_lp(4242, 512, 511)
    run_both(ctx, leq, tg, max_iter=120, tag="wander")


def test_resume_equals_single_run(ctx):
    leq, tg = H.gen_dense_lp(31337, 128, 127)
    sf = xp.slack_form(leq, tg)
    lp = ctx.large_lp(*sf[0].shape)
    lp.upload(*sf)
    o = H.slack_solve_oracle("f64", *sf, max_iter=9)
    assert o["status"] == xp.SIX_TIME_OUT
    assert lp.solve(4) == xp.SIX_TIME_OUT
    assert lp.solve(9) == xp.SIX_TIME_OUT
    a = lp.download(log_cap=64)
    assert a["iters"] == 9 and np.array_equal(a["log"], o["log"])
    assert np.array_equal(H.bits(a["tab"]), H.bits(o["tab"]))
    assert np.array_equal(H.bits(a["tgtf"]), H.bits(o["tgtf"]))
    lp.close()


def test_upload_leq_device_slack_form(ctx):
    leq, tg = H.gen_dense_lp(99, 40, 39)
    sf = xp.slack_form(leq, tg)
    lp = ctx.large_lp(*sf[0].shape)
    lp.upload_leq(leq, tg)
    st = lp.solve()
    a = lp.download(log_cap=4096)
    o = H.slack_solve_oracle("f64", *sf)
    assert st == o["status"] and a["iters"] == o["iters"]
    assert np.array_equal(H.bits(a["tab"]), H.bits(o["tab"]))
    lp.close()


def test_linearity_property_full_size_sample(ctx):
    """Size-independent property used at BASELINE sizes: K pivots leave every
    basic column a unit vector (up to round-off) and checksum is reproducible."""
    m, n = 1024, 1023
    lp = ctx.large_lp(m, n + m + 1)
    lp.fill_synthetic(7)
    lp.solve(20)
    c1 = lp.checksum()
    lp.fill_synthetic(7)
    lp.solve(20)
    c2 = lp.checksum()
    assert c1 == c2
    a = lp.download()
    tab = a["tab"]
    for r, bv in enumerate(a["eq2bv"]):
        col = tab[:, bv]
        assert abs(col[r] - 1.0) < 1e-9
        assert np.abs(np.delete(col, r)).max() < 1e-9
    lp.close()


def test_c3_full_size_properties_and_oracle_sample(ctx):
    """BASELINE config 3 at its full size (8192 x 16384, 1 GiB): (a) the state after K pivots
    does not depend on the block size -- k = 1 is the reference's own schedule of one tableau
    pass per pivot, k = 32 applies 32 pivots per pass -- checked through the position-keyed
    checksum of all 134 M doubles, the basis and the pivot sequence; (b) stopping and resuming
    gives the same bits as running through; (c) pivot sequence, basis and objective constant
    after a bounded sample equal the oracle's (the CPU needs ~0.15 s per pivot here)."""
    from xpoly_b200.synth import dense_lp
    m, n, K = 8192, 8191, 48
    lp = ctx.large_lp(m, n + m + 1)
    res = {}
    for k in (1, 32, 7):
        lp.set_block(k)
        lp.fill_synthetic(20261017)
        assert lp.solve(K) == xp.SIX_TIME_OUT
        a = lp.download(want_tab=False, log_cap=K)
        res[k] = (lp.checksum(), a["eq2bv"].copy(), a["log"].copy(), H.bits(a["tgtf"]).copy())
    for k in (32, 7):
        assert res[k][0] == res[1][0], ("checksum", k)
        assert np.array_equal(res[k][1], res[1][1]) and np.array_equal(res[k][2], res[1][2])
        assert np.array_equal(res[k][3], res[1][3])
    lp.set_block(0)
    lp.fill_synthetic(20261017)
    for stop in (5, 17, K):
        assert lp.solve(stop) == xp.SIX_TIME_OUT
    assert lp.checksum() == res[1][0]
    Ks = 10  # live oracle sample
    leq, tg = dense_lp(20261017, m, n)
    o = H.slack_solve_oracle("f64", *xp.slack_form(leq, tg), max_iter=Ks, log_cap=Ks)
    lp.fill_synthetic(20261017)
    lp.solve(Ks)
    a = lp.download(want_tab=False, log_cap=Ks)
    assert np.array_equal(a["log"], o["log"])
    assert np.array_equal(a["eq2bv"], o["eq2bv"])
    assert np.array_equal(H.bits(a["tgtf"]), H.bits(o["tgtf"]))
    lp.close()


def test_c3_full_size_checkpoints_vs_reference(ctx):
    """the 8192 x 16384 LP after K = 1, 10, 50, 200 pivots against the CPU side --
    position-keyed checksum of all 134 M tableau doubles, of the objective row, the basis and the
    whole pivot sequence.  tests/golden/c3_checkpoints.json holds what the oracle port AND the
    unmodified reference (TwoStageMethod, set_param(0, K)) computed for this seed
    (tools/c3_checkpoints.py; the two agree); the windowed and the full-width panel, blocked and
    unblocked, must all land on the same bits."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c3_checkpoints.json")))
    m, n = gold["m"], gold["n"]
    assert (m, n) == (8192, 8191) and gold["seed"] == 20261017
    for K in ("1", "10", "50", "200"):  # the fixture itself: reference == port
        for f in ("status", "tab", "tgtf", "eq2bv_sum"):
            assert gold["reference"][K][f] == gold["oracle"][K][f], (K, f)
    lp = ctx.large_lp(m, n + m + 1)
    wts = np.arange(1, m + 1, dtype=np.int64)
    for window, block in ((0, 0), (-1, 0), (0, 7)):
        lp.set_window(window)
        lp.set_block(block)
        lp.fill_synthetic(gold["seed"])
        assert (lp.window > 0) == (window == 0)
        for K in (1, 10, 50, 200):
            assert lp.solve(K) == xp.SIX_TIME_OUT
            ref = gold["reference"][str(K)]
            a = lp.download(want_tab=False, log_cap=K)
            ct, cg = lp.checksum()
            assert ct == ref["tab"], ("tableau checksum", window, block, K)
            assert cg == ref["tgtf"], ("objective checksum", window, block, K)
            assert int(a["eq2bv"].astype(np.int64).dot(wts)) == ref["eq2bv_sum"], ("basis", window, block, K)
            assert np.array_equal(a["log"], np.array(gold["oracle"]["200"]["log"])[:K]), ("pivot sequence", K)
    lp.close()


def _check_two_stage(ctx, leq, tg, K, tag):
    g = ctx.two_stage_f64_large(leq, tg, K)
    o = H.two_stage("oracle", "f64", leq, tg, K, want_log=True)
    m, n = leq.shape[0], leq.shape[1] - 1
    assert g["status"] == o["status"], (tag, g["status"], o["status"])
    assert g["pivots"] == len(o["log"]), (tag, "pivots", g["pivots"], len(o["log"]))
    if o["status"] == H.SIX_NO_PRI:
        return g
    Cc = n + m + 1
    assert o["cols"] == Cc
    assert np.array_equal(g["eq2bv"], o["eq2bv"]), (tag, "eq2bv")
    assert np.array_equal(H.bits(g["tgtf"]), H.bits(o["tgtf"])), (tag, "tgtf")
    assert np.array_equal(H.bits(g["maxv"]), H.bits(o["maxv"])), (tag, "maxv")
    assert np.array_equal(H.bits(g["slack_sol"]), H.bits(o["slack_sol"])), (tag, "sol")
    # the final tableau stays on the device: compare all of it through the position-keyed checksum
    import ctypes as C
    f = H.oracle().xo_checksum_f64
    f.restype = C.c_uint64
    tab = np.ascontiguousarray(o["tab"], dtype=np.float64)
    assert ctx.last_lp_checksum()[0] == int(f(H.P(tab), tab.shape[0], tab.shape[1])), (tag, "tableau checksum")
    return g


@pytest.mark.parametrize("m,n", [(3, 2), (40, 30), (130, 129), (150, 260), (300, 200)])
def test_two_stage_large_phase1_on_device(ctx, m, n):
    """TwoStageMethod on the HBM-resident path with constructBasicFeasibleSolution on the device
    (auxiliary column, forced first pivot, auxiliary solve, xa pivot-out, objective restoration,
    column deletion): bit for bit against the oracle, with and without phase 1, bounded and not."""
    seen = set()
    for k in range(6):
        leq, tg = H.gen_mixed_lp(31 * m + k, m, n, bneg=0.3 if k % 2 == 0 else 0.0)
        for K in ((H.NO_LIMIT, 7) if m <= 130 else (60, 400)):
            g = _check_two_stage(ctx, leq, tg, K, ("mixed", m, n, k, K))
            seen.add(g["status"])
    leq, tg = H.gen_dense_lp(5 * m, m, n)  # b > 0, c > 0: no auxiliary LP
    _check_two_stage(ctx, leq, tg, 40, ("dense", m, n))
    tg2 = -np.abs(tg)  # no positive cost: auxiliary LP although b > 0 (:1803)
    tg2[n] = 0.0
    _check_two_stage(ctx, leq, tg2, 40, ("nopos", m, n))
    assert len(seen) >= 1


def lower_bound_lp(seed, m, n, nneg):
    """A feasible LP that needs phase 1 and passes it: the dense family plus `nneg` lower bounds
    -x_j <= -l_j (negative constant terms), so the auxiliary LP reaches 0, xa is pivoted out (or
    already non-basic), the objective is restored by substitution and column xa is dropped."""
    r = np.random.RandomState(seed)
    leq = np.zeros((m + nneg, n + 1))
    leq[:m, :n] = r.uniform(0, 1, size=(m, n))
    leq[:m, n] = 1 + r.uniform(0, 1, size=m) * n
    for t in range(nneg):
        leq[m + t, t] = -1.0
        leq[m + t, n] = -0.01 * (t + 1)
    tg = np.zeros(n + 1)
    tg[:n] = r.uniform(0, 1, size=n)
    return leq, tg


@pytest.mark.parametrize("m,n,nneg", [(20, 15, 2), (100, 80, 3), (200, 150, 3), (300, 400, 4)])
def test_two_stage_large_phase1_succeeds(ctx, m, n, nneg):
    """The whole of constructBasicFeasibleSolution on the device with a successful phase 1 (the
    mixed-sign family above mostly ends in SIX_NO_PRI_FEASIBLE_SOL before the objective is
    restored): statuses past phase 1, every output bit for bit against the oracle."""
    seen = set()
    for s in range(5):
        leq, tg = lower_bound_lp(s, m, n, nneg)
        for K in ((H.NO_LIMIT, 9) if m <= 20 else (400,)):
            g = _check_two_stage(ctx, leq, tg, K, ("lb", m, n, s, K))
            seen.add(g["status"])
    assert H.SIX_NO_PRI not in seen and len(seen) >= 1


@pytest.mark.parametrize("env", [{"XP_FLUSH_NBUF": "2"}, {"XP_FLUSH_WIDE": "1"}, {"XP_FLUSH_WIDE": "1", "XP_FLUSH_NBUF": "2"}])
def test_flush_variants_bitwise(ctx, env, monkeypatch):
    """The alternative schedules of the tableau-update kernel -- two multiplier buffers with a block
    barrier per unit (the fallback where three buffers would cost the second CTA per SM) and the
    8 x 4 register tile (k_flush_w) -- against the oracle: whole tableau as uint64, k in {12, 16, 32}.
    The choice is cached per handle, so every case gets a fresh one."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for (m, n), seed in (((70, 130), 1), ((129, 64), 2), ((300, 215), 3)):
        leq, tg = H.gen_dense_lp(8300 + seed, m, n)
        sf = xp.slack_form(leq, tg)
        for k in (12, 16, 32):
            for K in (k, 3 * k + 5):
                lp = ctx.large_lp(*sf[0].shape)
                lp.set_block(k)
                lp.upload(*sf)
                st = lp.solve(K)
                g = lp.download(log_cap=1 << 16)
                g["status"] = st
                lp.close()
                assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=K), (env, m, n, k, K))
    # full-size cross-check of the variant against the default kernel (checksum of all 134 M doubles)
    m, n = 8192, 8191
    lp = ctx.large_lp(m, n + m + 1)
    lp.fill_synthetic(99)
    lp.solve(64)
    var = lp.checksum()
    lp.close()
    for k in env:
        monkeypatch.delenv(k)
    lp = ctx.large_lp(m, n + m + 1)
    lp.fill_synthetic(99)
    lp.solve(64)
    assert lp.checksum() == var
    lp.close()


@pytest.mark.parametrize("m,n,window,K,first", [(300, 1500, 512, 128, 0), (300, 1500, 256, 100, 0),
                                                (300, 1500, 1024, 64, 0), (520, 2100, 512, 120, 0),
                                                (520, 2101, 512, 120, 0), (257, 1300, 768, 97, 0),
                                                (300, 3500, 1024, 128, 512), (300, 3500, 768, 200, 256),
                                                (280, 4100, 1024, 256, 768)])
def test_two_stage_large_streamed_upload(ctx, m, n, window, K, first, monkeypatch):
    """xp_six_two_stage_f64_large uploading behind the solve: window columns first, the bounded
    solve runs on the early tiles while the rest of A is still on its way, the late pieces (one to
    three, `first` < window: the streamed run decides in a narrower window than the resident
    solve) replay the closed blocks out of the ring.  Bit for bit against the oracle, for runs the
    window decides alone and for runs a pricing scan leaves it (the full-width solve then
    continues)."""
    monkeypatch.setenv("XP_STREAM_MIN_MB", "0")
    if first:
        monkeypatch.setenv("XP_STREAM_FIRST", str(first))
    ctx.set_window(window)
    try:
        for seed in range(3):
            leq, tg = H.gen_dense_lp(8400 + seed, m, n)
            _check_two_stage(ctx, leq, tg, K, ("streamed", m, n, window, K, seed))
        leq, tg = H.gen_mixed_lp(8500, m, n)  # mixed signs: ratio-test failures inside the window run
        leq[:, n] = np.abs(leq[:, n])
        tg[:n] = np.abs(tg[:n])
        _check_two_stage(ctx, leq, tg, K, ("streamed-mixed", m, n, window, K))
    finally:
        ctx.set_window(0)
