"""GPU: the reference-side program of tests/adaptor/use_adaptor.cpp -- the reference's own
SIX<FloatMat,Float> / SIX<RMat,Rational> / MIP<RMat,Rational> call sites compiled against the
reference headers with xpoly_b200/host/xp_six.hpp -- runs on the device and reproduces the
reference's documented answer (example.cpp:89-93).  The binary is built by
tests/test_adaptor_cpu.py where /root/reference exists and travels in oracle/_ref/."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "use_adaptor")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/use_adaptor not built (no reference here)")
def test_reference_call_sites_run_on_the_gpu():
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "xpoly_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "status 0 max 2 x = (1.5555555555555556, 1.1111111111111112)" in r.stdout, r.stdout  # example.cpp:89-93
    exp = EXE + ".expected"  # the unmodified reference's answers, recorded at build time
    if os.path.exists(exp):
        assert r.stdout.strip().splitlines() == open(exp).read().strip().splitlines()
