"""torchrun worker for tests/test_sharded_gpu.py::test_multiprocess_ipc: every rank
solves its column shard, compares its slice and the replicated state with the
oracle bit for bit, and the shard checksums are summed across ranks."""
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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    one_gpu = bool(os.environ.get("XP_TEST_ONE_GPU"))
    # one-GPU mode: every process opens the SAME device; the exchange blocks still travel as
    # real cudaIpc handles between processes (kernels of different processes are time-sliced,
    # so every cross-rank wait costs a context switch: small cases only)
    local = 0 if one_gpu else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = xp.Context(local)
    cases = [("dense", 7001, 24, 23, H.NO_LIMIT, 0), ("dense", 7002, 64, 63, H.NO_LIMIT, 1 << 20),
             ("mixed", 3, 10, 9, H.NO_LIMIT, 0), ("mixed", 1, 12, 30, H.NO_LIMIT, 4),
             ("dense", 4242, 256, 255, 40, 0), ("dense", 4242, 256, 255, 70, 1 << 20)]
    if one_gpu:
        cases = [("dense", 7001, 24, 23, 12, -1), ("dense", 7002, 40, 39, 40, 1 << 20), ("mixed", 3, 10, 9, 30, 3)]
    for kind, seed, m, n, K, window in cases:
        if kind == "dense":
            leq, tg = H.gen_dense_lp(seed, m, n)
        else:
            leq, tg = H.gen_mixed_lp(seed, m, n)
            leq[:, n] = np.abs(leq[:, n])
        sf = xp.slack_form(leq, tg)
        lp = sharded.ShardedLP(ctx, m, sf[0].shape[1], rank, world, dist)
        lp.set_window(window)  # > 0: rank 0 decides runs of pivots alone (k_wpanel), peers replay
        lp.upload(*sf)
        st = lp.solve(K)
        g = lp.download(log_cap=1 << 16)
        o = H.slack_solve_oracle("f64", *sf, max_iter=K)
        sl = slice(lp.col0, lp.col0 + lp.local_cols)
        assert st == o["status"], (kind, seed, st, o["status"])
        assert g["iters"] == o["iters"]
        assert np.array_equal(g["log"], o["log"][: len(g["log"])])
        for k in ("eq2bv", "bv2eq", "nvset"):
            assert np.array_equal(g[k], o[k]), k
        assert np.array_equal(H.bits(g["tab"][:, sl]), H.bits(o["tab"][:, sl])), "tableau slice"
        assert np.array_equal(H.bits(g["tgtf"][sl]), H.bits(o["tgtf"][sl])), "objective slice"
        assert np.array_equal(H.bits(g["maxv"]), H.bits(o["maxv"]))
        assert np.array_equal(H.bits(g["sol"]), H.bits(o["sol"]))
        cs = torch.tensor([c % (1 << 63) for c in lp.checksum()], dtype=torch.int64)  # sum mod 2^63
        dist.all_reduce(cs)
        dist.barrier()
        lp.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED_WORKER_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
