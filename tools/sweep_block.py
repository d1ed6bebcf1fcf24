"""Dev tool: pivots/s of the c3 workload as a function of the block size k."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpoly_b200 as xp

m = int(os.environ.get("M", 8192))
n = m - 1
Cc = n + m + 1
P = int(os.environ.get("P", 192))
ctx = xp.Context(0)
lib = xp.lib()
lp = ctx.large_lp(m, Cc)
for k in [int(x) for x in os.environ.get("KS", "1,2,4,8,12,16,24,32").split(",")]:
    lp.set_block(k)
    lp.fill_synthetic(20261017)
    lp.solve(max(k, 1) * 2)          # warm-up
    lib.xp_lp_f64_profile(lp._h, 1)
    done = max(k, 1) * 2
    ms = 0.0
    reps = 3
    for _ in range(reps):
        st = lp.solve(done + P)
        done += P
        ms += ctx.last_kernel_ms
    nsw, sw, gap = C.c_uint64(0), C.c_double(0), C.c_double(0)
    lib.xp_lp_f64_profile_read(lp._h, C.byref(nsw), C.byref(sw), C.byref(gap))
    lib.xp_lp_f64_profile(lp._h, 0)
    piv = reps * P
    print(f"k={k:2d} status={st} {piv / (ms * 1e-3):9.1f} pivots/s  {ms * 1e3 / piv:7.2f} us/pivot  "
          f"flush avg {sw.value / max(nsw.value, 1) * 1e3:7.1f} us x{nsw.value}  "
          f"panel+gaps {gap.value * 1e3 / max(nsw.value - reps, 1) / max(k, 1):6.2f} us/pivot", flush=True)
lp.close()
ctx.close()
