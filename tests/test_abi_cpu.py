"""CPU: the C-ABI library loads, exports every symbol include/xpoly_b200.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "xpoly_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import xpoly_b200 as xp
    from xpoly_b200 import build
    build.build()
    lib = xp.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import xpoly_b200 as xp
    with pytest.raises(xp.XpolyError) as e:
        xp.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """Nothing under xpoly_b200/ or include/ may import, link or name the oracle."""
    bad = []
    for base in ("xpoly_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"xp_oracle|xo_[a-z]|libxpoly_ref|oracle/", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_status_codes_match_reference_values():
    import xpoly_b200 as xp
    assert (xp.SIX_SUCC, xp.SIX_UNBOUND, xp.SIX_NO_PRI_FEASIBLE_SOL, xp.SIX_OPTIMAL_IS_INFEASIBLE,
            xp.SIX_TIME_OUT) == (0, 1, 2, 3, 4)  # lpsol.h:198-202
    assert (xp.IP_SUCC, xp.IP_UNBOUND, xp.IP_NO_PRI_FEASIBLE_SOL,
            xp.IP_NO_BETTER_THAN_BEST_SOL) == (0, 1, 2, 3)  # lpsol.h:2082-2085
