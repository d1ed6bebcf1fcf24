"""Dev probe: c3-sized LP, blocked solve with and without the lookahead; XP_BLOCK_DBG timeline."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xpoly_b200 as xp

ctx = xp.Context(0)
m, n = 8192, 8191
modes = [("look", {})] + [(a, dict([a.split("=")])) for a in sys.argv[1:]] + [("nolook", {"XP_NO_LOOKAHEAD": "1"})] + [("nolook," + a, dict([a.split("=")])) for a in sys.argv[1:] if a.startswith("XP_FLUSH")]
for mode, env in modes:
    os.environ.update(env)
    lp = ctx.large_lp(m, n + m + 1)
    lp.fill_synthetic(2024)
    lp.solve(600)
    done = 600
    ms = 0.0
    for _ in range(10):
        done += 224
        lp.solve(done)
        ms += ctx.last_kernel_ms
    print(mode, "pivots/s %.0f" % (2240 / (ms * 1e-3)), "us per 32-pivot block %.1f" % (ms * 1e3 / 70), lp.checksum(), flush=True)
    lp.close()
