"""Column-sharded large LP, one process per GPU.

The pivot path has no host or library collective: the kernels exchange the
pricing candidate and the entering column through NVLink peer memory.  The only
thing that needs an out-of-band channel is the one-off all-gather of the 64-byte
CUDA IPC handles, done here with torch.distributed (any backend: the handles
are host bytes)."""
import numpy as np


LEADER_MIN = 3072  # columns rank 0 keeps at least (it leads the windowed runs); shard_lo() in xp_large_f64.cu


def shard_bounds(C, nranks):
    """Column range [lo, hi) of every rank, in units of two columns (the kernels use 128-bit
    accesses): an even split, except that rank 0 keeps at least LEADER_MIN columns when the even
    share would be smaller.  Mirrors shard_lo() in xp_large_f64.cu."""
    pairs = (C + 1) // 2
    if nranks > 2 and pairs // nranks < LEADER_MIN // 2 and pairs >= LEADER_MIN:
        rest = pairs - LEADER_MIN // 2
        lo = [0] + [min(C, LEADER_MIN + 2 * (rest * (r - 1) // (nranks - 1))) for r in range(1, nranks)] + [C]
    else:
        lo = [min(C, 2 * (pairs * r // nranks)) for r in range(nranks)] + [C]
    return [(lo[r], lo[r + 1]) for r in range(nranks)]


def owner_of(C, nranks, j):
    for r, (lo, hi) in enumerate(shard_bounds(C, nranks)):
        if lo <= j < hi:
            return r
    raise ValueError(j)


def allgather_handles(handle, dist, group=None):
    """All-gather one uint8[64] handle per rank; returns uint8 [world, 64]."""
    import torch
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(handle, dtype=np.uint8).copy())
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = t.to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy() for o in out])


class ShardedLP:
    """This rank's shard of an m x C tableau, attached to its peers."""

    def __init__(self, ctx, m, C, rank, world, dist):
        from . import LargeLP
        self.lp = LargeLP(ctx, m, C, rank, world)
        self.lp.peer_attach(allgather_handles(self.lp.peer_handle(), dist))
        dist.barrier()
        self._h = self.lp._h
        self.col0, self.local_cols = self.lp.col0, self.lp.local_cols

    def __getattr__(self, name):
        return getattr(self.lp, name)
