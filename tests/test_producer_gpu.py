"""GPU: the dependence queries the unmodified reference's producer path issued
(tests/golden/deppoly_queries.json, see test_producer_cpu.py) answered in one batch -- through the
C ABI and, where the replay binary was built, through XpHasSolutionBatch inside a program that runs
the reference's own DepPoly::is_empty loop -- must equal the reference's answers, query by query."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "deppoly_queries.json")
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_producer")


def test_recorded_producer_queries_one_batch(ctx, monkeypatch):
    d = json.load(open(GOLD))
    systems = [(np.array(q["leq"], dtype=np.int64).reshape(q["rows"], q["rhs_idx"] + 1), None) for q in d["queries"]]
    want = np.array([q["answer"] for q in d["queries"]], dtype=np.int32)
    got = ctx.has_solution_ragged(systems, is_int=True, is_unique=True)
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]
    monkeypatch.setenv("XP_HS_HOST", "1")  # the lock-step host path gives the same answers
    assert np.array_equal(ctx.has_solution_ragged(systems, is_int=True, is_unique=True), want)


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/ref_producer not built (no reference here)")
def test_reference_producer_loop_with_batched_answers():
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "xpoly_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([EXE, "replay"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches, 0 undecided" in r.stdout, r.stdout
