"""CPU: pins oracle/xp_oracle.c against tests/golden/reference_vectors.json, which
was produced by the UNMODIFIED reference (tests/golden/make_golden.py).  Bit-exact
for FP64 (hex floats), exact num/den for rationals."""
import json
import os

import numpy as np

import harness as H

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def test_example_float():
    ex = GOLD["example_float"]
    r = H.six_solve("oracle", "f64", 0, np.array(ex["leq"]), np.array(ex["tgtf"]))
    assert r["status"] == ex["status"] == 0
    assert np.array_equal(H.bits(r["v"]), H.bits(unhex(ex["v"])))
    assert np.array_equal(H.bits(r["sol"]), H.bits(unhex(ex["sol"])))
    assert r["v"][0] == 2.0  # example.cpp:89-93


def test_example_rational():
    ex = GOLD["example_rational"]
    leq, tg = H.to_rat(np.array(ex["leq"])), H.to_rat(np.array(ex["tgtf"]))
    mx = H.six_solve("oracle", "rat", 0, leq, tg)
    mn = H.six_solve("oracle", "rat", 1, leq, tg)
    assert mx["status"] == ex["max"]["status"] == H.SIX_UNBOUND          # example.cpp:163
    assert mn["status"] == 0 and mn["v"].tolist() == [23, 1]             # example.cpp:171-174
    assert mn["sol"].tolist() == ex["min"]["sol"]


def test_appendix_a4_a5_snapshots():
    for name in ("A4", "A5"):
        g = GOLD[name]
        leq, tg = H.gen_dense_lp(g["seed"], g["m"], g["n"])
        for s in g["snapshots"]:
            r = H.two_stage("oracle", "f64", leq, tg, s["K"])
            assert r["status"] == s["status"], (name, s["K"])
            assert r["eq2bv"].tolist() == s["eq2bv"], (name, s["K"])
            assert np.array_equal(H.bits(r["tgtf"][-1:]), H.bits(unhex(s["tgtf_rhs"])))
            assert np.array_equal(H.bits(r["maxv"]), H.bits(unhex(s["maxv"])))
            assert np.array_equal(H.bits(r["tab"][0][:8]), H.bits(unhex(s["tab_row0"])))
            assert np.array_equal(H.bits(np.array([r["tab"].sum()])), H.bits(unhex(s["tab_sum"])))
        e = H.six_solve("oracle", "f64", 0, leq, tg)
        assert e["status"] == g["maxm_status"]
        assert np.array_equal(H.bits(e["v"]), H.bits(unhex(g["maxm_v"])))
    # the published-in-SURVEY numbers themselves
    leq, tg = H.gen_dense_lp(12345, 256, 255)
    r = H.two_stage("oracle", "f64", leq, tg, want_log=True)
    assert r["status"] == 0 and len(r["log"]) == 14
    assert r["log"][:5, 1:].tolist() == [[0, 493, 238], [1, 408, 153], [2, 302, 47], [5, 1, 153],
                                         [29, 2, 47]]
    assert H.six_solve("oracle", "f64", 0, leq, tg)["v"][0] == 4.5154303644256268


def test_two_stage_f64_family():
    for d in GOLD["two_stage_f64"]:
        if d["gen"] == "dense":
            leq, tg = H.gen_dense_lp(d["seed"], d["m"], d["n"])
        else:
            leq, tg = H.gen_mixed_lp(d["seed"], d["m"], d["n"], bneg=0.3)
        r = H.two_stage("oracle", "f64", leq, tg)
        assert r["status"] == d["status"], d
        if d["status"] == 2:
            continue
        assert r["eq2bv"].tolist() == d["eq2bv"]
        for k in ("maxv", "tgtf", "slack_sol"):
            assert np.array_equal(H.bits(r[k]), H.bits(unhex(d[k]))), (d["seed"], k)
        assert np.array_equal(H.bits(np.array([r["tab"].sum()])), H.bits(unhex(d["tab_sum"])))


def test_two_stage_rat_family():
    for d in GOLD["two_stage_rat"]:
        leq, tg = H.gen_int_lp(d["seed"], d["m"], d["n"], **d["kw"])
        a0 = H.appro_count("oracle")
        r = H.two_stage("oracle", "rat", H.to_rat(leq), H.to_rat(tg))
        assert r["status"] == d["status"]
        assert H.appro_count("oracle") - a0 == d["appro"]
        assert r["eq2bv"].tolist() == d["eq2bv"]
        assert r["maxv"].tolist() == d["maxv"]
        assert r["tgtf"].tolist() == d["tgtf"]
        assert r["slack_sol"].tolist() == d["slack_sol"]


def test_mip_and_has_solution():
    for d in GOLD["mip_rat"]:
        leq, tg = H.gen_int_lp(d["seed"], d["m"], d["n"], alo=-1, ahi=4, density=0.8, blo=1, bhi=25)
        r = H.mip_solve("oracle", "rat", 0, 0, H.to_rat(leq), H.to_rat(tg))
        assert r["status"] == d["status"]
        if d["status"] == 0:
            assert r["v"].tolist() == d["v"] and r["sol"].tolist() == d["sol"]
        assert H.has_solution("oracle", H.to_rat(leq)) == d["has_solution"]
    A = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, 1], [-1, 1, -1]]
    B = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, -20]]
    Cc = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [2, -2, 1], [-2, 2, -1]]
    got = {nm: H.has_solution("oracle", H.to_rat(np.array(M, dtype=float)))
           for nm, M in (("A", A), ("B", B), ("C", Cc))}
    assert got == GOLD["has_solution_A6"] == {"A": 1, "B": 0, "C": 0}


def test_edge_cases():
    # 1x1 LPs, zero objective, zero rows
    for leq, tg in ((np.array([[1.0, 5.0]]), np.array([1.0, 0.0])),
                    (np.array([[0.0, 5.0]]), np.array([1.0, 0.0])),
                    (np.array([[1.0, -1.0]]), np.array([0.0, 0.0])),
                    (np.array([[-1.0, -1.0], [1.0, 3.0]]), np.array([-1.0, 2.0]))):
        r = H.two_stage("oracle", "f64", leq, tg)
        assert r["status"] in (0, 1, 2, 3)
        if H.ref() is not None:
            b = H.two_stage("ref", "f64", leq, tg)
            assert b["status"] == r["status"]
            if r["status"] != 2:
                assert np.array_equal(H.bits(b["tab"]), H.bits(r["tab"]))


def test_phase1_success_family():
    """TwoStageMethod where phase 1 succeeds (objective restored, column xa dropped)."""
    for k, d in enumerate(GOLD["two_stage_f64_phase1_ok"]):
        leq = unhex(d["leq"]).reshape(d["m"], d["n"] + 1)
        tg = unhex(d["tgtf"])
        r = H.two_stage("oracle", "f64", leq, tg)
        assert r["status"] == d["status"], k
        assert r["eq2bv"].tolist() == d["eq2bv"], k
        for a, b in (("maxv", "maxv"), ("tgtf", "tgtf_out"), ("slack_sol", "slack_sol")):
            assert np.array_equal(H.bits(r[a]), H.bits(unhex(d[b]))), (k, a)


def test_fea_schedule_and_has_solution_with_equalities():
    """The callers of SURVEY 8(f): FeaSchedule's equality-only MIPs and has_solution with
    equalities, as the unmodified reference answered them."""
    for k, d in enumerate(GOLD["fea_schedule_mip"]):
        eq, tg = np.array(d["eq"], dtype=np.int64), np.array(d["tgtf"], dtype=np.int64)
        for nm, is_min in (("max", 0), ("min", 1)):
            r = H.mip_solve("oracle", "rat", is_min, 0, None, H.to_rat(tg), eq=H.to_rat(eq))
            if d[nm] is None:
                assert r["status"] < 0, (k, nm)
                continue
            assert r["status"] == d[nm]["status"], (k, nm)
            if r["status"] == 0:
                assert r["v"].tolist() == d[nm]["v"] and r["sol"].tolist() == d[nm]["sol"], (k, nm)
    for k, d in enumerate(GOLD["has_solution_eq"]):
        leq, eq = np.array(d["leq"], dtype=np.int64), np.array(d["eq"], dtype=np.int64)
        assert H.has_solution("oracle", H.to_rat(leq), H.to_rat(eq)) == d["result"], k


def test_mip_rational_indicator_golden():
    """rational_indicator cases recorded from the unmodified reference (make_golden_ri.py)."""
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mip_indicator_vectors.json")))["cases"]
    assert len(cases) > 50
    for k, c in enumerate(cases):
        if c["is_bin"]:
            continue  # is_bin appends equalities and walks into convertEq2Ineq's column bug: reference-UB territory
        leq, tg = np.array(c["leq"], dtype=np.int64), np.array(c["tgtf"], dtype=np.int64)
        o = H.mip_solve_ri("oracle", c["is_min"], c["is_bin"], H.to_rat(leq), H.to_rat(tg), c["indicator"])
        assert o["status"] == c["status"], (k, o["status"], c["status"])
        assert o["v"].tolist() == c["v"], k
        if c["status"] == 0:
            assert o["sol"].tolist() == c["sol"], k
