"""Entry-level parity (SIX::maxm / minm, MIP::maxm / minm, Lineq::has_solution)
of the C ABI against the oracle: status, objective and solution -- bit-exact for
FP64, exact num/den for rationals (on LPs where the reference stayed exact)."""
import numpy as np
import pytest

import harness as H
import xpoly_b200 as xp

pytestmark = pytest.mark.gpu


def same_f64(g, o, tag):
    assert g["status"] == o["status"], (tag, g["status"], o["status"])
    assert np.array_equal(H.bits(g["v"]), H.bits(o["v"])), (tag, g["v"], o["v"])
    if o["status"] == 0:
        assert np.array_equal(H.bits(g["sol"]), H.bits(o["sol"])), tag


def same_rat(g, o, tag):
    assert g["status"] == o["status"], (tag, g["status"], o["status"])
    assert np.array_equal(g["v"], o["v"]), (tag, g["v"], o["v"])
    if o["status"] == 0:
        assert np.array_equal(g["sol"], o["sol"]), tag


def cover_lp(seed):
    r = np.random.RandomState(seed)
    m, n = r.randint(3, 10), r.randint(2, 8)
    A = r.randint(0, 4, size=(m, n)).astype(float)
    leq = np.zeros((m, n + 1))
    leq[:, :n] = -A
    leq[:, n] = -r.randint(1, 10, size=m)
    tg = np.zeros(n + 1)
    tg[:n] = r.randint(1, 6, size=n)
    return leq, tg


def test_example_float_golden(ctx):
    """src/example/example.cpp:54-93: max 2x1-x2 s.t. 2x1-x2<=2, x1-5x2<=-4 => 2 at (14/9, 10/9)."""
    leq = np.array([[2, -1, 2], [1, -5, -4]], dtype=float)
    tg = np.array([2, -1, 0], dtype=float)
    g = ctx.six_solve("f64", 0, leq, tg)
    assert g["status"] == 0 and g["v"][0] == 2.0
    assert g["sol"].tolist() == [1.5555555555555556, 1.1111111111111112, 1.0]


def test_example_rational_golden(ctx):
    """src/example/example.cpp:106-181: max unbounded, min = 23 at (10,5,3,2,3)."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    ex = gold["example_rational"]
    leq = np.array(ex["leq"], dtype=np.int64)
    tg = np.array(ex["tgtf"], dtype=np.int64)
    gmax = ctx.six_solve("rat", 0, leq, tg)
    gmin = ctx.six_solve("rat", 1, leq, tg)
    assert gmax["status"] == ex["max"]["status"] == xp.SIX_UNBOUND
    assert gmin["status"] == 0 and gmin["v"].tolist() == ex["min"]["v"] == [23, 1]
    assert gmin["sol"].tolist() == ex["min"]["sol"]


def test_maxm_minm_f64_vs_oracle(ctx):
    for seed in range(120):
        leq, tg = cover_lp(seed)
        for is_min in (0, 1):
            same_f64(ctx.six_solve("f64", is_min, leq, tg),
                     H.six_solve("oracle", "f64", is_min, leq, tg), ("cover", seed, is_min))
        leq, tg = H.gen_mixed_lp(seed, 9, 7, bneg=0.3)
        for is_min in (0, 1):
            same_f64(ctx.six_solve("f64", is_min, leq, tg),
                     H.six_solve("oracle", "f64", is_min, leq, tg), ("mixed", seed, is_min))


def test_maxm_minm_rat_vs_oracle(ctx):
    n_checked = 0
    for seed in range(150):
        for leq, tg in (cover_lp(seed), H.gen_int_lp(seed, 10, 8, alo=-2, ahi=3, density=0.6,
                                                      blo=-4, bhi=15)):
            for is_min in (0, 1):
                a0 = H.appro_count("oracle")
                o = H.six_solve("oracle", "rat", is_min, H.to_rat(leq), H.to_rat(tg))
                if H.appro_count("oracle") != a0:
                    continue
                same_rat(ctx.six_solve("rat", is_min, leq.astype(np.int64), tg.astype(np.int64)), o,
                         (seed, is_min))
                n_checked += 1
    assert n_checked > 500


def test_equalities_and_free_variables(ctx):
    """eq rows go through convertEq2Ineq (including its mis-indexed column,
    lpsol.h:1232); free variables (zero vc column) are split v = v' - v''."""
    r = np.random.RandomState(1)
    seen_ub = 0
    for seed in range(120):
        m, n = r.randint(3, 8), r.randint(2, 6)
        leq, tg = H.gen_int_lp(seed, m, n, alo=-2, ahi=3, density=0.7)
        k = r.randint(1, 3)
        E = np.zeros((k, n + 1))
        E[:, :n] = r.randint(-2, 3, size=(k, n))
        E[:, n] = r.randint(0, 6, size=k)
        o = H.six_solve("oracle", "f64", 0, leq, tg, None, E)
        g = ctx.six_solve("f64", 0, leq, tg, eq=E)
        if o["status"] == xp.ERR_REFERENCE_UB:
            seen_ub += 1
            assert g["status"] == xp.ERR_REFERENCE_UB
        else:
            same_f64(g, o, ("eq", seed))
        vc = np.zeros((n, n + 1))
        for i in range(n):
            if r.uniform() < 0.6:
                vc[i, i] = -1
        for is_min in (0, 1):
            same_f64(ctx.six_solve("f64", is_min, leq, tg, vc=vc),
                     H.six_solve("oracle", "f64", is_min, leq, tg, vc), ("free", seed, is_min))
    assert seen_ub > 0


def test_large_lp_through_entry_goes_to_hbm_path(ctx):
    """An LP too large for shared memory is routed to the HBM-resident path,
    including phase 1 on the device (negative right-hand sides)."""
    leq, tg = H.gen_dense_lp(3, 150, 149)
    same_f64(ctx.six_solve("f64", 0, leq, tg), H.six_solve("oracle", "f64", 0, leq, tg), "large")
    leq2 = leq.copy()
    leq2[::7, -1] = -0.5
    leq2[::7, :-1] *= -1.0
    same_f64(ctx.six_solve("f64", 0, leq2, tg, max_iter=60),
             H.six_solve("oracle", "f64", 0, leq2, tg, max_iter=60), "large-phase1")
    # phase 1 that succeeds (lower bounds), through maxm and -- on the explicit dual -- minm
    from test_large_f64_gpu import lower_bound_lp
    for s in range(3):
        l3, t3 = lower_bound_lp(s, 140, 120, 3)
        for is_min in (0, 1):
            same_f64(ctx.six_solve("f64", is_min, l3, t3, max_iter=300),
                     H.six_solve("oracle", "f64", is_min, l3, t3, max_iter=300), ("large-lb", s, is_min))


def test_batched_entry_f64_and_rat(ctx):
    lps = [H.gen_dense_lp(2024 + k, 16, 15) for k in range(64)]
    leq = np.stack([l for l, _ in lps])
    tg = np.stack([t for _, t in lps])
    for is_min in (0, 1):
        g = ctx.six_solve_batch("f64", is_min, leq, tg)
        for k, (l, t) in enumerate(lps):
            o = H.six_solve("oracle", "f64", is_min, l, t)
            same_f64(dict(status=g["status"][k], v=g["v"][k:k + 1], sol=g["sol"][k]), o, (k, is_min))
    lps = [cover_lp(900 + k) if k % 2 else H.gen_int_lp(k, 8, 6) for k in range(40)]
    lps = [(l, t) for l, t in lps if l.shape == lps[1][0].shape]
    leq = np.stack([l for l, _ in lps]).astype(np.int64)
    tg = np.stack([t for _, t in lps]).astype(np.int64)
    for is_min in (0, 1):
        g = ctx.six_solve_batch("rat", is_min, leq, tg)
        for k, (l, t) in enumerate(lps):
            o = H.six_solve("oracle", "rat", is_min, H.to_rat(l), H.to_rat(t))
            same_rat(dict(status=g["status"][k], v=g["v"][k], sol=g["sol"][k]), o, (k, is_min))


def knapsack(n, seed):
    r = np.random.RandomState(seed)
    w = r.randint(5, 41, size=n)
    p = r.randint(5, 61, size=n)
    leq = np.zeros((n + 1, n + 1), dtype=np.int64)
    leq[0, :n] = w
    leq[0, n] = int(w.sum() // 3)
    for j in range(n):
        leq[1 + j, j] = 1
        leq[1 + j, n] = 1
    tg = np.zeros(n + 1, dtype=np.int64)
    tg[:n] = p
    return leq, tg


def test_mip_rat_vs_oracle(ctx):
    for seed in range(80):
        r = np.random.RandomState(seed)
        m, n = r.randint(2, 7), r.randint(2, 6)
        leq, tg = H.gen_int_lp(seed, m, n, alo=-1, ahi=4, density=0.8, blo=1, bhi=25)
        for is_min in (0, 1):
            o = H.mip_solve("oracle", "rat", is_min, 0, H.to_rat(leq), H.to_rat(tg))
            g = ctx.mip_solve("rat", is_min, 0, leq.astype(np.int64), tg.astype(np.int64))
            same_rat(g, o, ("mip", seed, is_min))
            assert g["nodes"] == o["nodes"], ("nodes", seed, is_min)


def test_mip_f64_vs_oracle(ctx):
    for seed in range(40):
        r = np.random.RandomState(seed)
        m, n = r.randint(2, 7), r.randint(2, 6)
        leq, tg = H.gen_int_lp(seed, m, n, alo=-1, ahi=4, density=0.8, blo=1, bhi=25)
        o = H.mip_solve("oracle", "f64", 0, 0, leq, tg)
        same_f64(ctx.mip_solve("f64", 0, 0, leq, tg), o, ("mipf", seed))


def test_c5_knapsack_bnb(ctx):
    """Config 5 family.  n <= 50 solves; n = 200 has the reference answer
    IP_UNBOUND after one node."""
    for n, seed in ((10, 109), (30, 129), (50, 149)):
        leq, tg = knapsack(n, seed)
        o = H.mip_solve("oracle", "rat", 0, 0, H.to_rat(leq), H.to_rat(tg))
        g = ctx.mip_solve("rat", 0, 0, leq, tg)
        same_rat(g, o, ("knap", n))
        assert g["nodes"] == o["nodes"]
    leq, tg = knapsack(200, 99)
    g = ctx.mip_solve("rat", 0, 0, leq, tg)
    o = H.mip_solve("oracle", "rat", 0, 0, H.to_rat(leq), H.to_rat(tg))
    assert g["status"] == o["status"] == xp.IP_UNBOUND and g["nodes"] == o["nodes"] == 1


def test_mip_batch_lockstep_equals_individual(ctx):
    lps = [knapsack(12, 300 + k) for k in range(24)]
    leq = np.stack([l for l, _ in lps])
    tg = np.stack([t for _, t in lps])
    g = ctx.mip_solve_rat_batch(0, 0, leq, tg)
    for k, (l, t) in enumerate(lps):
        o = H.mip_solve("oracle", "rat", 0, 0, H.to_rat(l), H.to_rat(t))
        same_rat(dict(status=g["status"][k], v=g["v"][k], sol=g["sol"][k]), o, ("batch", k))
        assert g["nodes"][k] == o["nodes"]


def test_has_solution_dependence_queries(ctx):
    """SURVEY Appendix A6 (A true, B false, C false) + random systems vs the oracle."""
    A = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, 1], [-1, 1, -1]]
    Cc = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [2, -2, 1], [-2, 2, -1]]
    res = ctx.has_solution_batch(np.array([A, Cc], dtype=np.int64))
    assert res.tolist() == [1, 0]
    B = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, -20]]
    assert ctx.has_solution_batch(np.array([B], dtype=np.int64)).tolist() == [0]
    lps = [H.gen_int_lp(k, 6, 4, alo=-1, ahi=4, density=0.8, blo=1, bhi=25)[0] for k in range(100)]
    res = ctx.has_solution_batch(np.stack(lps).astype(np.int64))
    for k, l in enumerate(lps):
        assert res[k] == H.has_solution("oracle", H.to_rat(l)), k
    res2 = ctx.has_solution_batch(np.stack(lps).astype(np.int64), is_int=False, is_unique=False)
    for k, l in enumerate(lps):
        assert res2[k] == H.has_solution("oracle", H.to_rat(l), is_int=False, is_unique=False), k


def test_has_solution_ragged_with_equalities(ctx):
    """The producer's shape (poly.cpp:1166-1195): systems of different sizes, inequalities plus
    equalities, answered in one call; each against the oracle's Lineq::has_solution.  Systems on
    which the reference has undefined behaviour come back as XP_ERR_REFERENCE_UB."""
    r = np.random.RandomState(42)
    systems = []
    for k in range(160):
        n = int(r.randint(2, 6))
        m = int(r.randint(2, 9))
        leq = np.zeros((m, n + 1), dtype=np.int64)
        leq[:, :n] = r.randint(-2, 4, size=(m, n)) * (r.uniform(size=(m, n)) < 0.7)
        leq[:, n] = r.randint(0, 25, size=m)
        eq = None
        if k % 3 == 0:
            ke = int(r.randint(1, 3))
            eq = np.zeros((ke, n + 1), dtype=np.int64)
            eq[:, :n] = r.randint(-2, 3, size=(ke, n))
            eq[:, n] = r.randint(0, 10, size=ke)
        systems.append((leq, eq))
    systems.append(None)
    for is_int, is_unique in ((True, True), (False, False)):
        res = ctx.has_solution_ragged(systems[:-1], is_int=is_int, is_unique=is_unique)
        n_ub = 0
        for k, (leq, eq) in enumerate(systems[:-1]):
            o = H.has_solution("oracle", H.to_rat(leq), None if eq is None else H.to_rat(eq),
                               is_int=is_int, is_unique=is_unique)
            if res[k] < 0:
                assert res[k] == xp.ERR_REFERENCE_UB and eq is not None, (k, res[k])
                n_ub += 1
                continue
            assert res[k] == o, (k, res[k], o, is_int)
        assert n_ub < 40
    # the uniform entry point is the ragged one with equal shapes
    lps = [H.gen_int_lp(k, 6, 4, alo=-1, ahi=4, density=0.8, blo=1, bhi=25)[0] for k in range(50)]
    a = ctx.has_solution_batch(np.stack(lps).astype(np.int64))
    b = ctx.has_solution_ragged([(l.astype(np.int64), None) for l in lps])
    assert np.array_equal(a, b)
    # equalities without inequalities: the reference writes a 1 x 0 objective (linsys.cpp:851)
    e = np.array([[1, -1, 0]], dtype=np.int64)
    assert ctx.has_solution_ragged([(None, e)]).tolist() == [xp.ERR_REFERENCE_UB]


def fea_schedule_system(r):
    """The MIP PolyTran::FeaSchedule builds (poly.cpp:5094-5133): equalities only (Farkas
    identities over u- and lambda-multipliers), all variables >= 0, objective = sum of the u's."""
    nu, nl, ke = int(r.randint(2, 5)), int(r.randint(3, 7)), int(r.randint(2, 5))
    n = nu + nl
    eq = np.zeros((ke, n + 1), dtype=np.int64)
    eq[:, :n] = r.randint(-2, 3, size=(ke, n))
    eq[:, n] = r.randint(0, 6, size=ke)
    tg = np.zeros(n + 1, dtype=np.int64)
    tg[:nu] = 1
    return eq, tg


def test_fea_schedule_shape_mip(ctx):
    """maxm, then minm when that fails (poly.cpp:5126-5133), on equality-only
    integer programs, against the oracle: status, value, solution, node count."""
    r = np.random.RandomState(5)
    seen = set()
    for k in range(60):
        eq, tg = fea_schedule_system(r)
        for is_min in (0, 1):
            g = ctx.mip_solve("rat", is_min, 0, None, tg, eq=eq)
            o = H.mip_solve("oracle", "rat", is_min, 0, None, H.to_rat(tg), eq=H.to_rat(eq))
            if o["status"] < 0:  # the reference has undefined behaviour here: only the code is defined
                assert g["status"] == o["status"], (k, is_min, g["status"], o["status"])
                continue
            same_rat(g, o, ("fea", k, is_min))
            assert g["nodes"] == o["nodes"], (k, is_min)
            seen.add(o["status"])
            if o["status"] == 0:
                break
    assert {0, 1, 2} <= seen


def test_golden_reference_answers_for_the_8f_callers(ctx):
    """tests/golden (written by the unmodified reference): TwoStageMethod with a successful
    phase 1 on both FP64 paths, FeaSchedule-shape MIPs, has_solution with equalities."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    unhex = lambda xs: np.array([float.fromhex(x) for x in xs], dtype=np.float64)
    for k, d in enumerate(gold["two_stage_f64_phase1_ok"]):
        leq = unhex(d["leq"]).reshape(d["m"], d["n"] + 1)
        tg = unhex(d["tgtf"])
        b = ctx.two_stage_f64_batch(leq[None], tg[None])
        g = ctx.two_stage_f64_large(leq, tg)
        Cc = d["m"] + d["n"] + 1
        for nm, st, e2b, mv, tgo, ss in (("batched", b["status"][0], b["eq2bv"][0], b["maxv"][0:1], b["tgtf"][0], b["slack_sol"][0]),
                                         ("large", g["status"], g["eq2bv"], g["maxv"], g["tgtf"], g["slack_sol"])):
            assert st == d["status"], (k, nm)
            assert e2b[:d["m"]].tolist() == d["eq2bv"], (k, nm)
            assert np.array_equal(H.bits(mv), H.bits(unhex(d["maxv"]))), (k, nm)
            assert np.array_equal(H.bits(tgo[:Cc]), H.bits(unhex(d["tgtf_out"]))), (k, nm)
            assert np.array_equal(H.bits(ss[:Cc]), H.bits(unhex(d["slack_sol"]))), (k, nm)
    for k, d in enumerate(gold["fea_schedule_mip"]):
        eq, tg = np.array(d["eq"], dtype=np.int64), np.array(d["tgtf"], dtype=np.int64)
        for nm, is_min in (("max", 0), ("min", 1)):
            g = ctx.mip_solve("rat", is_min, 0, None, tg, eq=eq)
            if d[nm] is None:
                assert g["status"] == xp.ERR_REFERENCE_UB, (k, nm)
                continue
            assert g["status"] == d[nm]["status"], (k, nm)
            if g["status"] == 0:
                assert g["v"].tolist() == d[nm]["v"] and g["sol"].tolist() == d[nm]["sol"], (k, nm)
    systems = [(np.array(d["leq"], dtype=np.int64), np.array(d["eq"], dtype=np.int64)) for d in gold["has_solution_eq"]]
    res = ctx.has_solution_ragged(systems)
    assert res.tolist() == [d["result"] for d in gold["has_solution_eq"]]


def test_mip_rational_indicator(ctx):
    """MIP with rational_indicator (lpsol.h:2369-2391) through xp_mip_solve_rat_ri: the cases the
    unmodified reference answered (tests/golden/mip_indicator_vectors.json) and live oracle cases."""
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mip_indicator_vectors.json")))["cases"]
    n_ok = 0
    for k, c in enumerate(cases):
        if c["is_bin"]:
            continue
        leq, tg = np.array(c["leq"], dtype=np.int64), np.array(c["tgtf"], dtype=np.int64)
        g = ctx.mip_solve_rat_ri(c["is_min"], 0, leq, tg, c["indicator"])
        assert g["status"] == c["status"], (k, g["status"], c["status"])
        assert g["v"].tolist() == c["v"], k
        if c["status"] == 0:
            assert g["sol"].tolist() == c["sol"], k
            n_ok += 1
    assert n_ok > 10
    rs = np.random.RandomState(3)
    for seed in range(30):
        leq, tg = H.gen_int_lp(2900 + seed, 6, 5, alo=-1, ahi=4, density=0.7, blo=1, bhi=17)
        ind = (rs.uniform(size=6) < 0.4).astype(np.uint8)
        a0 = H.appro_count("oracle")
        o = H.mip_solve_ri("oracle", 0, 0, H.to_rat(leq), H.to_rat(tg), ind)
        if H.appro_count("oracle") != a0:
            continue
        g = ctx.mip_solve_rat_ri(0, 0, leq.astype(np.int64), tg.astype(np.int64), ind)
        assert g["status"] == o["status"] and g["nodes"] == o["nodes"], seed
        assert np.array_equal(g["v"], o["v"]), seed


def test_general_variable_constraints(ctx):
    """vc beyond -x <= 0 (other diagonals, constant terms, free variables): only the feasibility
    check of the optimal exit reads it (lpsol.h:798-802).  FP64 goes through the HBM-resident path
    with the two vectors on the device, the exact path re-decides the verdict on the solution row."""
    r = np.random.RandomState(77)
    seen = set()
    for seed in range(120):
        m, n = r.randint(3, 8), r.randint(2, 6)
        leq, tg = H.gen_int_lp(3000 + seed, m, n, alo=-1, ahi=3, density=0.7, blo=0, bhi=15)
        vc = np.zeros((n, n + 1))
        for i in range(n):
            u = r.uniform()
            if u < 0.2:
                continue
            vc[i, i] = -r.randint(1, 4)
            if u > 0.5:
                vc[i, n] = r.randint(-6, 4)
        for is_min in (0, 1):
            same_f64(ctx.six_solve("f64", is_min, leq, tg, vc=vc),
                     H.six_solve("oracle", "f64", is_min, leq, tg, vc), ("vc-f64", seed, is_min))
            a0 = H.appro_count("oracle")
            o = H.six_solve("oracle", "rat", is_min, H.to_rat(leq), H.to_rat(tg), H.to_rat(vc))
            if H.appro_count("oracle") != a0:
                continue
            g = ctx.six_solve("rat", is_min, leq.astype(np.int64), tg.astype(np.int64), vc=vc.astype(np.int64))
            same_rat(g, o, ("vc-rat", seed, is_min))
            seen.add((is_min, o["status"]))
    assert (0, 0) in seen and (0, 3) in seen


def test_minm_large_dual_built_on_device(ctx, monkeypatch):
    """SIX::minm of LPs beyond shared memory: the explicit dual (calcDualMaxm) is built on the device
    from the caller's primal (tiled transposition, -A^T | c, objective -b) -- same bits as the oracle
    and as the host-built dual (XP_HOST_DUAL=1), for bounded covering problems (phase 1 on the dual:
    its constant column c has no negative entry but the dual objective -b has no positive one) and for
    dense LPs."""
    r = np.random.RandomState(12)
    seen = set()
    for seed, (m, n) in enumerate([(150, 140), (97, 260), (300, 120), (180, 181)]):
        A = r.randint(0, 4, size=(m, n)).astype(float) * (r.uniform(size=(m, n)) < 0.4)
        leq = np.zeros((m, n + 1))
        leq[:, :n] = -A
        leq[:, n] = -r.randint(1, 10, size=m)       # A x >= b
        tg = np.zeros(n + 1)
        tg[:n] = r.randint(1, 9, size=n)            # min c x
        o = H.six_solve("oracle", "f64", 1, leq, tg)
        g = ctx.six_solve("f64", 1, leq, tg)
        same_f64(g, o, ("cover", m, n))
        seen.add(o["status"])
        leq2, tg2 = H.gen_dense_lp(9300 + seed, m, n)
        same_f64(ctx.six_solve("f64", 1, leq2, tg2, max_iter=60), H.six_solve("oracle", "f64", 1, leq2, tg2, max_iter=60),
                 ("dense-min", m, n))
        monkeypatch.setenv("XP_HOST_DUAL", "1")
        h = ctx.six_solve("f64", 1, leq, tg)
        monkeypatch.delenv("XP_HOST_DUAL")
        same_f64(g, h, ("host-vs-device dual", m, n))
        assert np.array_equal(g["eq2bv"][: n], h["eq2bv"][: n])
    # (sparse covering LPs of this size end SIX_UNBOUND in the reference -- tabu exhaustion, SURVEY
    # App. B 5b; min problems that succeed go the same way in test_large_lp_through_entry_goes_to_hbm_path)
    from test_large_f64_gpu import lower_bound_lp
    for s_ in range(3):
        l3, t3 = lower_bound_lp(10 + s_, 140, 120, 3)
        o = H.six_solve("oracle", "f64", 1, l3, t3, max_iter=300)
        same_f64(ctx.six_solve("f64", 1, l3, t3, max_iter=300), o, ("lb-min", s_))
        seen.add(o["status"])
    assert len(seen) >= 1


def test_has_solution_device_resident_vs_lockstep_and_oracle(ctx, monkeypatch):
    """Lineq::has_solution with the whole query on the device (one warp per query: both MIPs, DFS
    branch & bound, every node relaxation in registers) against the lock-step host path
    (XP_HS_HOST=1) on 6 000 random dependence systems, and against the oracle on a sample; all
    four (is_int_sol, is_unique_sol) combinations."""
    r = np.random.RandomState(4711)
    systems = []
    for k in range(6000):
        nq, mq = int(r.randint(1, 7)), int(r.randint(1, 16))  # up to 15 rows: both kernel sizes and, beyond 24 rows at the deepest node, the host path
        sysm = np.zeros((mq, nq + 1), dtype=np.int64)
        sysm[:, :nq] = r.randint(-3, 5, size=(mq, nq)) * (r.uniform(size=(mq, nq)) < 0.7)
        sysm[:, nq] = r.randint(-4, 25, size=mq)
        systems.append((sysm, None))
    for is_int in (True, False):
        for is_unique in (True, False):
            dev = ctx.has_solution_ragged(systems, is_int=is_int, is_unique=is_unique)
            monkeypatch.setenv("XP_HS_HOST", "1")
            host = ctx.has_solution_ragged(systems, is_int=is_int, is_unique=is_unique)
            monkeypatch.delenv("XP_HS_HOST")
            bad = np.nonzero(dev != host)[0]
            assert len(bad) == 0, (is_int, is_unique, bad[:10], dev[bad[:10]], host[bad[:10]])
            for k in range(0, 6000, 11):
                a0 = H.appro_count("oracle")
                o = H.has_solution("oracle", H.to_rat(systems[k][0]), is_int=is_int, is_unique=is_unique)
                if H.appro_count("oracle") == a0:
                    assert dev[k] == o, (is_int, is_unique, k, dev[k], o)
            assert set(np.unique(dev).tolist()) >= {0, 1}
