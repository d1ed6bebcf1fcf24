"""Dev tool: time the c4 exact batch (10k LPs of tableau 24x48) through the host-pointer call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpoly_b200 as xp
ctx = xp.Context(0)
r = np.random.RandomState(777)
B, m, n = int(os.environ.get("B", 10000)), 24, 23
A = r.randint(0, 4, size=(B, m, n)) * (r.uniform(size=(B, m, n)) < 0.3)
leq = np.zeros((B, m, n + 1), dtype=np.int64); leq[:, :, :n] = A; leq[:, :, n] = r.randint(0, 21, size=(B, m))
tg = np.zeros((B, n + 1), dtype=np.int64); tg[:, :n] = r.randint(1, 6, size=(B, n))
ctx.two_stage_i64_batch(leq[:64], tg[:64])
for _ in range(int(os.environ.get("REPS", 3))):
    res = ctx.two_stage_i64_batch(leq, tg)
print(f"kernel {ctx.last_kernel_ms:.2f} ms  pivots {int(res['pivots'].sum())}  max {int(res['pivots'].max())}")
