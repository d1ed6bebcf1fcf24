"""CPU: the producer side of the small-LP batches.  oracle/ref_producer.cpp drives the
UNMODIFIED reference's DepPoly::is_empty (poly.cpp:530-573: Lineq::reduce pre-filter, then
Lineq::has_solution) on 400 dependence polyhedra and records -- through ld --wrap, no reference
source touched -- every system that reaches has_solution with the reference's answer.  Here:
the committed fixture tests/golden/deppoly_queries.json is what that program writes (where the
reference is present), the oracle agrees with every recorded answer, and the replay binary that
answers the same queries in ONE XpHasSolutionBatch (xp_six.hpp -> libxpoly_b200.so) links."""
import json
import os
import subprocess

import numpy as np
import pytest

import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "deppoly_queries.json")
REFENG = "/root/reference/src/eng"
REFCOM = "/root/reference/src/com"


def test_oracle_agrees_with_recorded_reference_answers():
    d = json.load(open(GOLD))
    assert len(d["queries"]) > 300 and 0 < d["empty"] < d["polyhedra"]
    seen = set()
    for k, q in enumerate(d["queries"]):
        leq = np.array(q["leq"], dtype=np.int64).reshape(q["rows"], q["rhs_idx"] + 1)
        a0 = H.appro_count("oracle")
        o = H.has_solution("oracle", H.to_rat(leq), is_int=bool(q["is_int"]), is_unique=bool(q["is_unique"]))
        if H.appro_count("oracle") != a0:
            continue
        assert o == q["answer"], k
        seen.add(o)
    assert seen == {0, 1}


@pytest.mark.skipif(not os.path.exists(os.path.join(REFENG, "poly.cpp")), reason="reference sources not present")
def test_fixture_is_what_the_unmodified_reference_produces(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "producer"])
    out = tmp_path / "q.json"
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_producer_rec"), "record", str(out)],
                          stdout=subprocess.DEVNULL)
    assert json.load(open(out)) == json.load(open(GOLD))
    # the replay binary: same driver + the adaptor header, answered by libxpoly_b200.so (run on the GPU box)
    from xpoly_b200 import build
    build.build()
    ref_objs = [os.path.join(ROOT, "oracle", "_ref", f"{n}.o")
                for n in ("sgraph", "smempool", "comf", "strbuf", "bs", "rational", "flty", "linsys", "xmat", "ltype")]
    eng_objs = [os.path.join(ROOT, "oracle", "_ref", f"eng_{n}.o") for n in ("poly", "ldtran", "depvecs")]
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_producer")
    cmd = ["g++", "-D_LINUX_", "-DXP_WITH_ADAPTOR", "-Wno-write-strings", "-O2", "-w", "-I", REFCOM, "-I", REFENG,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "xpoly_b200", "host"),
           os.path.join(ROOT, "oracle", "ref_producer.cpp"), *eng_objs, *ref_objs,
           "-Wl,--wrap=_ZN4xcom5Lineq12has_solutionERKNS_4RMatES3_RS1_jbb",
           "-L", os.path.join(ROOT, "xpoly_b200"), "-lxpoly_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "xpoly_b200"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
