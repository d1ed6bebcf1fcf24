"""Generates tests/golden/mip_indicator_vectors.json by running the UNMODIFIED reference
(oracle/_ref/libxpoly_ref.so): MIP<RMat,Rational>::maxm / minm with rational_indicator
(lpsol.h:2626-2657) on seeded small integer programs.  Run where /root/reference exists."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness as H  # noqa: E402

out = []
rs = np.random.RandomState(11)
for seed in range(60):
    m, n = [(4, 3), (6, 4), (7, 5), (5, 6)][seed % 4]
    leq, tg = H.gen_int_lp(1900 + seed, m, n, alo=-1, ahi=4, density=0.7, blo=1, bhi=17)
    ind = (rs.uniform(size=n + 1) < 0.5).astype(np.uint8)
    for is_min in (0, 1):
        for is_bin in (0, 1):
            if is_bin and is_min:
                continue
            a0 = H.appro_count("ref")
            r = H.mip_solve_ri("ref", is_min, is_bin, H.to_rat(leq), H.to_rat(tg), ind)
            if H.appro_count("ref") != a0:
                continue
            out.append(dict(leq=leq.astype(int).tolist(), tgtf=tg.astype(int).tolist(), indicator=ind.tolist(),
                            is_min=is_min, is_bin=is_bin, status=int(r["status"]), v=r["v"].tolist(),
                            sol=r["sol"].tolist() if r["status"] == 0 else None))
json.dump(dict(generator="tests/golden/make_golden_ri.py", cases=out),
          open(os.path.join(HERE, "mip_indicator_vectors.json"), "w"))
print(len(out), "cases", sum(c["status"] == 0 for c in out), "IP_SUCC")
