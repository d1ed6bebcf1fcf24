"""Generates tests/golden/c3_checkpoints.json: state of the config-3 LP (8192 x 16384 tableau,
synthetic seed 20261017) after K in {1, 10, 50, 200} simplex iterations, as computed on the CPU
by (a) the oracle port's solveSlackForm and (b) -- where oracle/_ref exists -- the UNMODIFIED
reference's TwoStageMethod.  Stored per K: the position-keyed checksum of the whole tableau
(xo_checksum_f64, the key function of xp_lp_f64_checksum), the checksum of the objective row,
eq2bv, and the pivot log.  TEST INFRASTRUCTURE; run once here (the reference needs ~10 GiB and
~0.5 s per pivot at this size), the GPU test compares against the committed file.

  python tools/c3_checkpoints.py [--ks 1,10,50,200] [--no-ref]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import xpoly_b200 as xp  # noqa: E402
from xpoly_b200.synth import dense_lp  # noqa: E402

SEED, M, N = 20261017, 8192, 8191


def cks(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    a2 = a.reshape(1, -1) if a.ndim == 1 else a
    f = H.oracle().xo_checksum_f64
    f.restype = C.c_uint64
    return int(f(H.P(a2), a2.shape[0], a2.shape[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ks", default="1,10,50,200")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--m", type=int, default=M)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "c3_checkpoints.json"))
    a = ap.parse_args()
    m, n = a.m, a.m - 1
    ks = [int(k) for k in a.ks.split(",")]
    leq, tg = dense_lp(SEED, m, n)
    out = {"seed": SEED, "m": m, "n": n, "C": n + m + 1, "generator": "tools/c3_checkpoints.py",
           "key": "sum mix64(bits ^ mix64(i*C+j)) mod 2^64 (splitmix64 finaliser)", "oracle": {}, "reference": {}}
    sf = xp.slack_form(leq, tg)
    for K in ks:
        t0 = time.time()
        o = H.slack_solve_oracle("f64", *sf, max_iter=K, log_cap=max(K, 1))
        out["oracle"][str(K)] = {"status": int(o["status"]), "tab": cks(o["tab"]), "tgtf": cks(o["tgtf"]),
                                 "eq2bv_sum": int(o["eq2bv"].astype(np.int64).dot(np.arange(1, m + 1))),
                                 "log": o["log"].tolist() if K <= 200 else None}
        print("oracle", K, time.time() - t0, flush=True)
        json.dump(out, open(a.out, "w"))
    if not a.no_ref and H.ref() is not None:
        for K in ks:
            t0 = time.time()
            r = H.two_stage("ref", "f64", leq, tg, K)
            out["reference"][str(K)] = {"status": int(r["status"]), "tab": cks(r["tab"]), "tgtf": cks(r["tgtf"]),
                                        "eq2bv_sum": int(r["eq2bv"].astype(np.int64).dot(np.arange(1, m + 1)))}
            print("reference", K, time.time() - t0, flush=True)
            json.dump(out, open(a.out, "w"))
    json.dump(out, open(a.out, "w"), indent=0)


if __name__ == "__main__":
    main()
