#!/usr/bin/env python
"""Regenerates tests/golden/reference_vectors.json by running the UNMODIFIED
reference (oracle/_ref/libxpoly_ref.so, built from /root/reference by
oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Contents: the reference's own worked examples (src/example/example.cpp:52-181),
the seeded instances of SURVEY.md Appendix A4/A5, state snapshots after K pivots
(TwoStageMethod + set_param(0,K), the parity hook of SURVEY 8c), exact-rational
solves, MIP / has_solution answers.  Floats are stored as hex strings (bit
exact), rationals as [num, den]."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    assert H.ref() is not None, "oracle/_ref/libxpoly_ref.so missing: make -C oracle ref"
    out = {}
    # --- example.cpp:54-93 (Float)
    leq = np.array([[2, -1, 2], [1, -5, -4]], dtype=float)
    tg = np.array([2, -1, 0], dtype=float)
    r = H.six_solve("ref", "f64", 0, leq, tg)
    out["example_float"] = dict(leq=leq.tolist(), tgtf=tg.tolist(), status=r["status"],
                                v=hx(r["v"]), sol=hx(r["sol"]))
    # --- example.cpp:106-181 (Rational): 8 constraints, 5 variables
    L = [[-1, 0, 0, 0, 0, -10], [-1, -1, 0, 0, 0, -8], [-1, -1, -1, 0, 0, -9],
         [-1, -1, -1, -1, 0, -11], [0, -1, -1, -1, -1, -13], [0, 0, -1, -1, -1, -8],
         [0, 0, 0, -1, -1, -5], [0, 0, 0, 0, -1, -3]]  # example.cpp:138-146
    T = [1, 1, 1, 1, 1, 0]                              # example.cpp:135-136
    leq_r, tg_r = np.array(L, dtype=float), np.array(T, dtype=float)
    mx = H.six_solve("ref", "rat", 0, H.to_rat(leq_r), H.to_rat(tg_r))
    mn = H.six_solve("ref", "rat", 1, H.to_rat(leq_r), H.to_rat(tg_r))
    out["example_rational"] = dict(
        leq=L, tgtf=T, max=dict(status=mx["status"], v=mx["v"].tolist()),
        min=dict(status=mn["status"], v=mn["v"].tolist(), sol=mn["sol"].tolist()))
    # --- Appendix A4 / A5 seeded instances (std::mt19937_64 stream)
    for name, (seed, m, n, Ks) in dict(A4=(12345, 256, 255, [1, 5, 14, H.NO_LIMIT]),
                                       A5=(12345, 8, 7, [2, H.NO_LIMIT])).items():
        leq, tg = H.gen_dense_lp(seed, m, n)
        snaps = []
        for K in Ks:
            r = H.two_stage("ref", "f64", leq, tg, K)
            snaps.append(dict(K=K, status=r["status"], eq2bv=r["eq2bv"].tolist(),
                              tgtf_rhs=hx(r["tgtf"][-1:]), maxv=hx(r["maxv"]),
                              tab_row0=hx(r["tab"][0][:8]), tab_sum=hx([r["tab"].sum()])))
        e = H.six_solve("ref", "f64", 0, leq, tg)
        out[name] = dict(seed=seed, m=m, n=n, snapshots=snaps, maxm_status=e["status"],
                         maxm_v=hx(e["v"]))
    # --- small seeded families: full state after TwoStageMethod
    fam = []
    for k in range(40):
        m, n = [(6, 5), (8, 7), (10, 9), (16, 15)][k % 4]
        leq, tg = (H.gen_mixed_lp(k, m, n, bneg=0.3) if k % 2 else H.gen_dense_lp(1000 + k, m, n))
        r = H.two_stage("ref", "f64", leq, tg)
        d = dict(gen="mixed_bneg0.3" if k % 2 else "dense", seed=(k if k % 2 else 1000 + k), m=m,
                 n=n, status=r["status"])
        if r["status"] != 2:
            d.update(eq2bv=r["eq2bv"].tolist(), maxv=hx(r["maxv"]), tgtf=hx(r["tgtf"]),
                     slack_sol=hx(r["slack_sol"]), tab_sum=hx([r["tab"].sum()]))
        fam.append(d)
    out["two_stage_f64"] = fam
    # --- exact rational family (c4 shape and a mixed-sign 8x7 family)
    rat = []
    for k in range(40):
        if k % 2:
            m, n, kw = 8, 7, dict(alo=-3, ahi=3, density=0.5)
        else:
            m, n, kw = 24, 23, dict()
        leq, tg = H.gen_int_lp(k, m, n, **kw)
        a0 = H.appro_count("ref")
        r = H.two_stage("ref", "rat", H.to_rat(leq), H.to_rat(tg))
        d = dict(seed=k, m=m, n=n, kw=kw, status=r["status"],
                 appro=int(H.appro_count("ref") - a0), eq2bv=r["eq2bv"].tolist(),
                 maxv=r["maxv"].tolist(), tgtf=r["tgtf"].tolist(),
                 slack_sol=r["slack_sol"].tolist())
        rat.append(d)
    out["two_stage_rat"] = rat
    # --- MIP + has_solution
    mips = []
    for k in range(30):
        r0 = np.random.RandomState(k)
        m, n = r0.randint(2, 7), r0.randint(2, 6)
        leq, tg = H.gen_int_lp(k, m, n, alo=-1, ahi=4, density=0.8, blo=1, bhi=25)
        r = H.mip_solve("ref", "rat", 0, 0, H.to_rat(leq), H.to_rat(tg))
        mips.append(dict(seed=k, m=int(m), n=int(n), status=r["status"], v=r["v"].tolist(),
                         sol=r["sol"].tolist() if r["status"] == 0 else None,
                         has_solution=int(H.has_solution("ref", H.to_rat(leq)))))
    out["mip_rat"] = mips
    A = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, 1], [-1, 1, -1]]
    B = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, -20]]
    Cc = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [2, -2, 1], [-2, 2, -1]]
    out["has_solution_A6"] = {nm: int(H.has_solution("ref", H.to_rat(np.array(M, dtype=float))))
                              for nm, M in (("A", A), ("B", B), ("C", Cc))}
    # --- phase 1 that SUCCEEDS (lower bounds -x_j <= -l_j): aux optimum 0, xa pivoted out,
    #     objective restored by substitution, column xa dropped, main solve
    ok1 = []
    for k in range(24):
        m, n, nneg = [(5, 4, 1), (12, 9, 2), (20, 15, 2)][k % 3]
        r0 = np.random.RandomState(500 + k)
        leq = np.zeros((m + nneg, n + 1))
        leq[:m, :n] = r0.uniform(0, 1, size=(m, n))
        leq[:m, n] = 1 + r0.uniform(0, 1, size=m) * n
        for t in range(nneg):
            leq[m + t, t] = -1.0
            leq[m + t, n] = -0.01 * (t + 1)
        tg = np.zeros(n + 1)
        tg[:n] = r0.uniform(0, 1, size=n)
        r = H.two_stage("ref", "f64", leq, tg)
        assert r["status"] != 2
        ok1.append(dict(leq=hx(leq), tgtf=hx(tg), m=m + nneg, n=n, status=r["status"],
                        eq2bv=r["eq2bv"].tolist(), maxv=hx(r["maxv"]), tgtf_out=hx(r["tgtf"]),
                        slack_sol=hx(r["slack_sol"])))
    out["two_stage_f64_phase1_ok"] = ok1
    # --- PolyTran::FeaSchedule's MIP shape (poly.cpp:5094-5133): equalities only, objective =
    #     sum of the first nu variables, maxm then minm
    fea = []
    r0 = np.random.RandomState(5)
    for k in range(40):
        nu, nl, ke = int(r0.randint(2, 5)), int(r0.randint(3, 7)), int(r0.randint(2, 5))
        n = nu + nl
        eq = np.zeros((ke, n + 1), dtype=np.int64)
        eq[:, :n] = r0.randint(-2, 3, size=(ke, n))
        eq[:, n] = r0.randint(0, 6, size=ke)
        tg = np.zeros(n + 1, dtype=np.int64)
        tg[:nu] = 1
        d = dict(eq=eq.tolist(), tgtf=tg.tolist())
        for nm, is_min in (("max", 0), ("min", 1)):
            if H.mip_solve("oracle", "rat", is_min, 0, None, H.to_rat(tg), eq=H.to_rat(eq))["status"] < 0:
                d[nm] = None  # the reference has undefined behaviour here (lpsol.h:1232)
                continue
            r = H.mip_solve("ref", "rat", is_min, 0, None, H.to_rat(tg), eq=H.to_rat(eq))
            d[nm] = dict(status=r["status"], v=r["v"].tolist(),
                         sol=r["sol"].tolist() if r["status"] == 0 else None)
        fea.append(d)
    out["fea_schedule_mip"] = fea
    # --- Lineq::has_solution with equalities, on systems free of the reference's UB
    hse = []
    r0 = np.random.RandomState(77)
    for k in range(150):
        n, m = int(r0.randint(2, 5)), int(r0.randint(2, 7))
        leq = np.zeros((m, n + 1), dtype=np.int64)
        leq[:, :n] = r0.randint(-2, 4, size=(m, n))
        leq[:, n] = r0.randint(0, 20, size=m)
        eq = np.zeros((1, n + 1), dtype=np.int64)
        eq[:, :n] = r0.randint(-2, 3, size=(1, n))
        eq[:, n] = r0.randint(0, 8)
        tg = np.zeros(n + 1, dtype=np.int64)
        tg[:n] = ((leq[:, :n] != 0).any(axis=0) | (eq[:, :n] != 0).any(axis=0)).astype(np.int64)
        amax = H.mip_solve("oracle", "rat", 0, 0, H.to_rat(leq), H.to_rat(tg), eq=H.to_rat(eq))
        if amax["status"] < 0:
            continue
        if amax["status"] != 0 and \
                H.mip_solve("oracle", "rat", 1, 0, H.to_rat(leq), H.to_rat(tg), eq=H.to_rat(eq))["status"] < 0:
            continue
        hse.append(dict(leq=leq.tolist(), eq=eq.tolist(),
                        result=int(H.has_solution("ref", H.to_rat(leq), H.to_rat(eq)))))
    out["has_solution_eq"] = hse
    json.dump(out, open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=0)
    print("wrote reference_vectors.json:", {k: (len(v) if isinstance(v, list) else "obj")
                                            for k, v in out.items()})


if __name__ == "__main__":
    main()
