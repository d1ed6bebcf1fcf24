"""CPU: live differential test of the oracle against the compiled reference
(oracle/_ref).  Skipped where /root/reference was never available."""
import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.skipif(H.ref() is None, reason="oracle/_ref not built (no reference here)")


def eqv(kind, x, y):
    return np.array_equal(H.bits(x), H.bits(y)) if kind == "f64" else np.array_equal(x, y)


def same_state(kind, leq, tg, K=H.NO_LIMIT):
    a = H.two_stage("oracle", kind, leq, tg, K)
    b = H.two_stage("ref", kind, leq, tg, K)
    assert (a["status"], a["rows"], a["cols"], a["rhs_idx"]) == (b["status"], b["rows"], b["cols"],
                                                                 b["rhs_idx"])
    if a["status"] != 2:
        for k in ("tab", "tgtf", "maxv", "slack_sol"):
            assert eqv(kind, a[k], b[k]), k
        for k in ("eq2bv", "bv2eq", "nvset", "bvset"):
            assert np.array_equal(a[k], b[k]), k
    return a["status"]


def test_two_stage_f64_state_bit_identical():
    seen = set()
    for seed in range(60):
        for m, n in ((6, 5), (10, 9), (16, 15)):
            seen.add(same_state("f64", *H.gen_dense_lp(1000 + seed, m, n)))
            seen.add(same_state("f64", *H.gen_mixed_lp(seed, m, n)))
            leq, tg = H.gen_mixed_lp(seed, m, n, bneg=0.3)
            seen.add(same_state("f64", leq, tg))
            for K in (0, 1, 3):
                same_state("f64", leq, tg, K)
    assert seen == {0, 1, 2, 3}


def test_two_stage_rat_state_and_appro_count():
    for seed in range(80):
        for m, n, kw in ((8, 7, dict(alo=-3, ahi=3, density=0.5)), (12, 11, dict())):
            leq, tg = H.gen_int_lp(seed, m, n, **kw)
            a0, b0 = H.appro_count("oracle"), H.appro_count("ref")
            same_state("rat", H.to_rat(leq), H.to_rat(tg))
            assert H.appro_count("oracle") - a0 == H.appro_count("ref") - b0
    for seed in range(12):  # c4 shape: the lossy appro() path shows up here
        leq, tg = H.gen_int_lp(seed, 24, 23)
        a0, b0 = H.appro_count("oracle"), H.appro_count("ref")
        same_state("rat", H.to_rat(leq), H.to_rat(tg))
        assert H.appro_count("oracle") - a0 == H.appro_count("ref") - b0


def test_maxm_minm_mip_has_solution():
    for seed in range(60):
        r = np.random.RandomState(seed)
        m, n = r.randint(3, 9), r.randint(2, 7)
        A = r.randint(0, 4, size=(m, n)).astype(float)
        leq = np.zeros((m, n + 1))
        leq[:, :n] = -A
        leq[:, n] = -r.randint(1, 10, size=m)
        tg = np.zeros(n + 1)
        tg[:n] = r.randint(1, 6, size=n)
        for kind, L, T in (("f64", leq, tg), ("rat", H.to_rat(leq), H.to_rat(tg))):
            for is_min in (0, 1):
                a = H.six_solve("oracle", kind, is_min, L, T)
                b = H.six_solve("ref", kind, is_min, L, T)
                assert a["status"] == b["status"] and eqv(kind, a["v"], b["v"])
                if a["status"] == 0:
                    assert eqv(kind, a["sol"], b["sol"])
        leq2, tg2 = H.gen_int_lp(seed, m, n, alo=-1, ahi=4, density=0.8, blo=1, bhi=25)
        a = H.mip_solve("oracle", "rat", 0, 0, H.to_rat(leq2), H.to_rat(tg2))
        b = H.mip_solve("ref", "rat", 0, 0, H.to_rat(leq2), H.to_rat(tg2))
        assert a["status"] == b["status"]
        if a["status"] == 0:
            assert np.array_equal(a["v"], b["v"]) and np.array_equal(a["sol"], b["sol"])
        assert H.has_solution("oracle", H.to_rat(leq2)) == H.has_solution("ref", H.to_rat(leq2))


def test_equalities_including_reference_bug_path():
    r = np.random.RandomState(7)
    n_ub = 0
    for seed in range(80):
        m, n = r.randint(3, 8), r.randint(2, 6)
        leq, tg = H.gen_int_lp(seed, m, n, alo=-2, ahi=3, density=0.7)
        k = r.randint(1, 3)
        E = np.zeros((k, n + 1))
        E[:, :n] = r.randint(-2, 3, size=(k, n))
        E[:, n] = r.randint(0, 6, size=k)
        a = H.six_solve("oracle", "f64", 0, leq, tg, None, E)
        if a["status"] == -100:  # the reference would read out of bounds (lpsol.h:1232)
            n_ub += 1
            continue
        b = H.six_solve("ref", "f64", 0, leq, tg, None, E)
        assert a["status"] == b["status"] and eqv("f64", a["v"], b["v"])
    assert n_ub > 0


def test_fea_schedule_shape_mip_and_has_solution_with_equalities():
    """The two callers of section 8(f): PolyTran::FeaSchedule's equality-only integer programs
    (poly.cpp:5094-5133: maxm, then minm) and Lineq::has_solution on systems with equalities."""
    r = np.random.RandomState(5)
    for k in range(40):
        nu, nl, ke = int(r.randint(2, 5)), int(r.randint(3, 7)), int(r.randint(2, 5))
        n = nu + nl
        eq = np.zeros((ke, n + 1), dtype=np.int64)
        eq[:, :n] = r.randint(-2, 3, size=(ke, n))
        eq[:, n] = r.randint(0, 6, size=ke)
        tg = np.zeros(n + 1, dtype=np.int64)
        tg[:nu] = 1
        for is_min in (0, 1):
            a = H.mip_solve("oracle", "rat", is_min, 0, None, H.to_rat(tg), eq=H.to_rat(eq))
            if a["status"] < 0:
                continue  # the reference has undefined behaviour here
            b = H.mip_solve("ref", "rat", is_min, 0, None, H.to_rat(tg), eq=H.to_rat(eq))
            assert a["status"] == b["status"], (k, is_min)
            if a["status"] == 0:
                assert np.array_equal(a["v"], b["v"]) and np.array_equal(a["sol"], b["sol"])
    checked = 0
    for k in range(120):
        n, m = int(r.randint(2, 5)), int(r.randint(2, 7))
        leq = np.zeros((m, n + 1), dtype=np.int64)
        leq[:, :n] = r.randint(-2, 4, size=(m, n))
        leq[:, n] = r.randint(0, 20, size=m)
        eq = np.zeros((1, n + 1), dtype=np.int64)
        eq[:, :n] = r.randint(-2, 3, size=(1, n))
        eq[:, n] = r.randint(0, 8)
        # has_solution = MIP max, then MIP min, on the all-ones objective restricted to the columns
        # that occur (reviseTargetFunc).  Where one of them walks into convertEq2Ineq's
        # out-of-bounds read (lpsol.h:1232, possible at any B&B node) the reference's answer is
        # whatever the heap held; the oracle reports XO_ERR_REFERENCE_UB and the case is skipped.
        tg = np.zeros(n + 1, dtype=np.int64)
        tg[:n] = ((leq[:, :n] != 0).any(axis=0) | (eq[:, :n] != 0).any(axis=0)).astype(np.int64)
        amax = H.mip_solve("oracle", "rat", 0, 0, H.to_rat(leq), H.to_rat(tg), eq=H.to_rat(eq))
        if amax["status"] < 0:
            continue
        if amax["status"] != 0:
            amin = H.mip_solve("oracle", "rat", 1, 0, H.to_rat(leq), H.to_rat(tg), eq=H.to_rat(eq))
            if amin["status"] < 0:
                continue
        checked += 1
        assert H.has_solution("oracle", H.to_rat(leq), H.to_rat(eq)) == \
            H.has_solution("ref", H.to_rat(leq), H.to_rat(eq)), k
    assert checked >= 15


def test_mip_rational_indicator():
    """MIP::is_satisfying with rational_indicator (lpsol.h:2369-2391): marked entries may stay
    rational and are never branched on -- status, value and solution against the unmodified
    reference, general-integer and 0-1, max and min."""
    rs = np.random.RandomState(5)
    seen = set()
    for seed in range(40):
        m, n = [(4, 3), (6, 4), (7, 5)][seed % 3]
        leq, tg = H.gen_int_lp(900 + seed, m, n, alo=-1, ahi=4, density=0.7, blo=1, bhi=17)
        ind = (rs.uniform(size=n + 1) < 0.5).astype(np.uint8)
        for is_min in (0, 1):
            a0 = H.appro_count("ref")
            b = H.mip_solve_ri("ref", is_min, 0, H.to_rat(leq), H.to_rat(tg), ind)
            if H.appro_count("ref") != a0:
                continue
            a = H.mip_solve_ri("oracle", is_min, 0, H.to_rat(leq), H.to_rat(tg), ind)
            assert a["status"] == b["status"], (seed, is_min, a["status"], b["status"])
            assert np.array_equal(a["v"], b["v"]), (seed, is_min)
            if a["status"] == 0:
                assert np.array_equal(a["sol"], b["sol"]), (seed, is_min)
                seen.add(bool((a["sol"][:, 1] != 1).any()))  # some accepted solutions really are fractional
    assert seen == {False, True}


def general_vc(r, n):
    """Variable constraints beyond -x <= 0: other negative diagonals and non-zero constant terms
    (lower bounds x >= c/d).  No free variables here: the reference records the split of a free
    variable with the variadic INTMat::sete (lpsol.h:1377), which is broken on x86-64 (SURVEY 8c
    caveat 2) -- the compiled reference crashes on them; the oracle-only tests cover that path."""
    vc = np.zeros((n, n + 1))
    for i in range(n):
        u = r.uniform()
        vc[i, i] = -r.randint(1, 4)
        if u > 0.5:
            vc[i, n] = r.randint(-6, 4)
    return vc


def test_general_variable_constraints():
    """is_feasible reads vc(i,i) and vc(i,rhs) (lpsol.h:798-802): the final verdict for general
    variable constraints, FP64 and exact, max and min."""
    r = np.random.RandomState(77)
    seen = set()
    for seed in range(150):
        m, n = r.randint(3, 8), r.randint(2, 6)
        leq, tg = H.gen_int_lp(3000 + seed, m, n, alo=-1, ahi=3, density=0.7, blo=0, bhi=15)
        vc = general_vc(r, n)
        for is_min in (0, 1):
            a = H.six_solve("oracle", "f64", is_min, leq, tg, vc)
            b = H.six_solve("ref", "f64", is_min, leq, tg, vc)
            assert a["status"] == b["status"] and eqv("f64", a["v"], b["v"]), (seed, is_min, "f64")
            a0 = H.appro_count("ref")
            b = H.six_solve("ref", "rat", is_min, H.to_rat(leq), H.to_rat(tg), H.to_rat(vc))
            if H.appro_count("ref") == a0:
                a = H.six_solve("oracle", "rat", is_min, H.to_rat(leq), H.to_rat(tg), H.to_rat(vc))
                assert a["status"] == b["status"] and np.array_equal(a["v"], b["v"]), (seed, is_min, "rat")
                seen.add((is_min, a["status"]))
    assert (0, 0) in seen and (0, 3) in seen
