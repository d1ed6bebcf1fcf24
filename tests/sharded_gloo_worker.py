"""torchrun/gloo worker (CPU only) for tests/test_sharded_gloo_cpu.py.

Executable model of the column-sharded pivot loop of xp_large_f64.cu: every rank
holds a column slice of the tableau and of the objective row plus replicas of the constant
column, the basis maps and the pair-tabu table; per pivot the ranks exchange exactly (1) one
pricing candidate each (all-reduce MIN, lowest index wins) and (2) the entering column from its
owner (broadcast).  The merged result must equal the oracle bit for bit, which is the property
the CUDA kernels rely on.  Also checks the host-side handshake helpers of xpoly_b200/sharded.py.
TEST INFRASTRUCTURE: numpy restatement, never used by the product."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
import xpoly_b200 as xp  # noqa: E402
from xpoly_b200 import sharded  # noqa: E402

BIG = 0x7FFFFFFF
EPS = 1e-17


def feq(a, b):  # Float::operator==, flty.cpp:41-58
    if (a > 0 and b < 0) or (a < 0 and b > 0):
        return False
    a, b = abs(a), abs(b)
    if (a == 0.0 and b <= EPS) or (b == 0.0 and a <= EPS):
        return True
    return abs(a - b) <= EPS


def allmin(v):
    t = torch.tensor([v], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


def allor(v):
    t = torch.tensor([int(v)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item())


def solve_sharded(sf, rank, world, max_iter):
    tab, tg, nvset, _, bv2eq, eq2bv = [np.array(a, copy=True) for a in sf]
    m, C = tab.shape
    n = C - 1
    lo, hi = sharded.shard_bounds(C, world)[rank]
    T, tgl = tab[:, lo:hi].copy(), tg[lo:hi].copy()       # my slices
    rhs, tg_rhs = tab[:, n].copy(), tg[n]                 # replicas
    tabu = np.zeros((n, n), dtype=bool)
    log = []
    cnt = 0

    def eligible(j):  # canBeNVCandidate, lpsol.h:124
        return (~tabu[j]).sum() - (0 if tabu[j, j] else 1) > 0

    def can_bv(j):    # canBeBVCandidate, lpsol.h:140
        return (~tabu[:, j]).sum() - (0 if tabu[j, j] else 1) > 0

    while cnt < max_iter:
        while True:
            best, anypos = BIG, False
            for j in range(lo, min(hi, n)):
                if nvset[j] and tgl[j - lo] > 0.0:
                    anypos = True
                    if eligible(j):
                        best = j
                        break
            q, anypos = allmin(best), allor(anypos)           # exchange (1)
            zl = n if q == BIG else q
            for j in range(lo, min(hi, zl)):
                if not nvset[j]:
                    tgl[j - lo] = 0.0
            if q == BIG:
                assert not anypos, "pair-search fallback not modelled"
                return 0, T, tgl, rhs, tg_rhs, eq2bv, nvset, log, cnt, (lo, hi)
            owner = sharded.owner_of(C, world, q)
            col = torch.zeros(m + 1, dtype=torch.float64)
            if owner == rank:
                col[:m] = torch.from_numpy(T[:, q - lo].copy())
                col[m] = tgl[q - lo]
            dist.broadcast(col, src=owner)                     # exchange (2)
            col = col.numpy()
            p = -1
            for pas in (0, 1):                                 # findPivotBV, lpsol.h:552-663
                bestv = None
                for i in range(m):
                    a = col[i]
                    if feq(a, 0.0) or (pas == 0 and not a > 0.0):
                        continue
                    bv = eq2bv[i]
                    if tabu[q, bv] or not can_bv(bv):
                        continue
                    v = rhs[i] / a
                    if bestv is None or v < bestv:
                        bestv, p = v, i
                if p >= 0:
                    break
            if p >= 0:
                break
            tabu[q, :] = True                                  # disableNV, lpsol.h:114
            tabu[q, q] = False
        bv = eq2bv[p]
        tabu[q, bv] = True                                     # genPair
        log.append((q, bv, p))
        pv, cq = col[p], col[m]
        r = 1.0 / pv
        r1, r0 = feq(r, 1.0), feq(r, 0.0)
        c1, c0 = feq(cq, 1.0), feq(cq, 0.0)
        def sc(x):  # Matrix::mulOfRow short-circuits, matt.h:1358-1367
            return x if r1 else (np.zeros_like(x) if r0 else x * r)
        prow = sc(T[p].copy())
        prow_rhs = float(sc(np.float64(rhs[p])))
        f = -col[:m]
        for i in range(m):
            if i != p:
                T[i] = T[i] + f[i] * prow
                rhs[i] = rhs[i] + f[i] * prow_rhs
        T[p], rhs[p] = prow, prow_rhs
        t = prow * -1.0
        if hi == C:
            t[n - lo] = -t[n - lo]
        t = np.zeros_like(t) if c0 else (t if c1 else t * cq)  # Matrix::mul short-circuits, matt.h:1335-1341
        tgl = t + tgl
        tt = -(prow_rhs * -1.0)
        tt = 0.0 if c0 else (tt if c1 else tt * cq)
        tg_rhs = tt + tg_rhs
        nvset[q], nvset[bv] = 0, 1
        eq2bv[p] = q
        bv2eq[q], bv2eq[bv] = p, -1
        cnt += 1
    return 4, T, tgl, rhs, tg_rhs, eq2bv, nvset, log, cnt, (lo, hi)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # --- handshake helpers
    mine = np.full(xp.PEER_HANDLE_BYTES, rank + 1, dtype=np.uint8)
    allh = sharded.allgather_handles(mine, dist)
    assert allh.shape == (world, xp.PEER_HANDLE_BYTES)
    assert all((allh[r] == r + 1).all() for r in range(world))
    for C in (10, 18, 19, 64):
        b = sharded.shard_bounds(C, world)
        assert b[0][0] == 0 and b[-1][1] == C
        assert all(sharded.owner_of(C, world, j) == r for r, (lo, hi) in enumerate(b) for j in range(lo, hi))
    # --- the sharded pivot loop against the oracle
    for seed, (m, n, K) in enumerate([(8, 7, 3), (12, 11, 6), (16, 15, 10), (9, 9, 5)]):
        leq, tg = H.gen_dense_lp(8100 + seed, m, n)
        sf = xp.slack_form(leq, tg)
        st, T, tgl, rhs, tg_rhs, eq2bv, nvset, log, cnt, (lo, hi) = solve_sharded(sf, rank, world, K)
        o = H.slack_solve_oracle("f64", *sf, max_iter=K)
        if st == 4:
            assert o["status"] == 4 and o["iters"] == cnt
        assert [tuple(x) for x in o["log"][:len(log)]] == log, (seed, "pivot sequence")
        assert np.array_equal(H.bits(T), H.bits(o["tab"][:, lo:hi])), (seed, "tableau slice")
        assert np.array_equal(H.bits(tgl), H.bits(o["tgtf"][lo:hi])), (seed, "objective slice")
        assert np.array_equal(H.bits(rhs), H.bits(o["tab"][:, -1])), (seed, "constant column replica")
        assert np.array_equal(H.bits(np.array([tg_rhs])), H.bits(o["tgtf"][-1:])), (seed, "objective constant replica")
        assert np.array_equal(eq2bv, o["eq2bv"]) and np.array_equal(nvset, o["nvset"])
    dist.barrier()
    if rank == 0:
        print("SHARDED_GLOO_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
