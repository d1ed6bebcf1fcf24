"""CPU: xpoly_b200/host/xp_six.hpp (the adaptor INTEGRATION.md describes) compiles against the
reference's own headers with the reference's flags and links against libxpoly_b200.so plus the
reference's objects.  Skipped where /root/reference is absent (the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFCOM = "/root/reference/src/com"


@pytest.mark.skipif(not os.path.exists(os.path.join(REFCOM, "lpsol.h")), reason="reference sources not present")
def test_adaptor_compiles_and_links():
    import xpoly_b200
    from xpoly_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    objs = [os.path.join(ROOT, "oracle", "_ref", f"{n}.o")
            for n in ("sgraph", "smempool", "comf", "strbuf", "bs", "rational", "flty", "linsys", "xmat", "ltype")]
    exe = os.path.join(ROOT, "oracle", "_ref", "use_adaptor")  # travels to the GPU box (tests/test_adaptor_gpu.py runs it)
    cmd = ["g++", "-D_LINUX_", "-Wno-write-strings", "-O2", "-w", "-I", REFCOM,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "xpoly_b200", "host"),
           os.path.join(ROOT, "tests", "adaptor", "use_adaptor.cpp"), *objs,
           "-L", os.path.join(ROOT, "xpoly_b200"), "-lxpoly_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "xpoly_b200"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    # the specialisations really replaced the template bodies: the binary imports the C ABI
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for s in ("xp_six_maxm_f64", "xp_six_minm_rat", "xp_mip_solve_rat", "xp_has_solution_rat_ragged",
              "xp_six_two_stage_f64_large_vc", "xp_ctx_last_lp_download", "xp_ctx_create"):
        assert s in syms, s
    # what the UNMODIFIED reference answers for the same three calls (the GPU test compares)
    import numpy as np
    import harness as H
    leq = np.array([[2, -1, 2], [1, -5, -4]], dtype=float)
    tg = np.array([2, -1, 0], dtype=float)
    a = H.six_solve("ref", "f64", 0, leq, tg)
    b = H.six_solve("ref", "rat", 1, H.to_rat(leq), H.to_rat(tg))
    c = H.mip_solve("ref", "rat", 0, 0, H.to_rat(leq), H.to_rat(tg))
    t = H.two_stage("ref", "f64", leq, tg)  # the unmodified TwoStageMethod on the same LP (phase 1 included)
    g17 = lambda xs: "".join(" %.17g" % x for x in xs)
    lines = ["status %d max %.17g x = (%.17g, %.17g)" % (a["status"], a["v"][0], a["sol"][0], a["sol"][1]),
             "minm_rat status %d v %d/%d" % (b["status"], b["v"][0], b["v"][1]),
             "mip_max_rat status %d v %d/%d x = (%d/%d, %d/%d)" % (c["status"], c["v"][0], c["v"][1],
                                                                  *c["sol"][0], *c["sol"][1])]
    lines[1:1] = ["two_stage status %d maxv %.17g rhs_idx %d eq2bv %d %d bv2eq %d %d %d %d nv %d%d%d%d"
                  % (t["status"], t["maxv"][0], t["rhs_idx"], *t["eq2bv"][:2], *t["bv2eq"][:4], *t["nvset"][:4]),
                  "two_stage tableau %d x %d:%s" % (t["rows"], t["cols"], g17(t["tab"].ravel())),
                  "two_stage tgtf:" + g17(t["tgtf"]), "two_stage slack_sol:" + g17(t["slack_sol"]),
                  "two_stage vc 4 x 5 diag: -1 -1 -1 -1"]
    qa = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, 1], [-1, 1, -1]]
    qb = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [1, -1, -20]]
    qc = [[-1, 0, -1], [1, 0, 10], [0, -1, -1], [0, 1, 10], [2, -2, 1], [-2, 2, -1]]
    hs = [H.has_solution("ref", H.to_rat(np.array(q, dtype=np.int64))) for q in (qa, qb, qc)]
    lines.append("has_solution %d %d %d" % tuple(hs))
    open(os.path.join(ROOT, "oracle", "_ref", "use_adaptor.expected"), "w").write("\n".join(lines) + "\n")
