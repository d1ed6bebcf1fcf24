"""CPU, world_size 2 (and 3) over gloo: the column-sharding logic of the large-LP path --
shard bounds / ownership, the IPC-handle all-gather, and an executable numpy model of the
per-pivot exchange (candidate all-reduce MIN + entering-column broadcast) whose merged result
must equal the oracle bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pivot_loop_over_gloo(world):
    port = 29710 + world
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_GLOO_OK" in r.stdout
