import numpy as np, time, sys
sys.path.insert(0, ".")
import xpoly_b200 as xp
ctx = xp.Context(0)
r = np.random.RandomState(777)
T, nk = 256, 40
w = r.randint(5, 41, size=(T, nk)); pr = r.randint(5, 61, size=(T, nk))
L = np.zeros((T, nk + 1, nk + 1), dtype=np.int64)
L[:, 0, :nk] = w; L[:, 0, nk] = w.sum(axis=1) // 3
for j in range(nk):
    L[:, 1 + j, j] = 1; L[:, 1 + j, nk] = 1
G = np.zeros((T, nk + 1), dtype=np.int64); G[:, :nk] = pr
ctx.mip_solve_rat_batch(0, 0, L[:4], G[:4])
l0 = ctx.launches
t0 = time.perf_counter(); res = ctx.mip_solve_rat_batch(0, 0, L, G); dt = time.perf_counter() - t0
print("wall ms", dt * 1e3, "nodes", int(res["nodes"].sum()), "launches", ctx.launches - l0)
