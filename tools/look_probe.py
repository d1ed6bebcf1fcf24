"""Dev probe: c3-sized LP, blocked solve under different settings (environment variables as arguments)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import xpoly_b200 as xp

ctx = xp.Context(0)
m, n = 8192, 8191
modes = [("default", {})] + [(a, dict([a.split("=")])) for a in sys.argv[1:]]
for mode, env in modes:
    os.environ.update(env)
    lp = ctx.large_lp(m, n + m + 1)
    lp.fill_synthetic(2024)
    lp.solve(600)
    done = 600
    ms = 0.0
    ws = []
    for _ in range(20):
        done += 224
        lp.solve(done)
        ms += ctx.last_kernel_ms
        ws.append(lp.window)
    import ctypes as C
    out = (C.c_longlong * 16)()
    xp.lib().xp_lp_f64_debug_state(lp._h, out)
    print("   pivots outside k_wpanel:", out[1] - out[13], "launches", ctx.launches)
    print(mode, "pivots/s %.0f" % (4480 / (ms * 1e-3)), "us per 32-pivot block %.1f" % (ms * 1e3 / 140), lp.checksum(), "windows", ws[::3], flush=True)
    lp.close()
    for k in env:
        del os.environ[k]
