"""Column-sharded HBM path: G shards of one LP exchange the pricing
candidate and the entering column through peer memory inside the kernels.  The
merged result must be bit-identical to the oracle (and hence to one GPU).

* in-process: G shards on cuda:0, one host thread + stream per shard
  (xp_lp_f64_peer_attach_local) -- runs on a single-GPU box;
* multi-process: torchrun, one process per GPU, CUDA IPC handles all-gathered
  with torch.distributed -- skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import harness as H
import xpoly_b200 as xp
from xpoly_b200 import sharded

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def solve_sharded_inproc(G, sf, max_iters, device=0, block=0, window=0):
    """Returns one merged result dict per entry of max_iters (successive resumes)."""
    tab, tg = sf[0], sf[1]
    m, Cc = tab.shape
    ctxs = [xp.Context(device) for _ in range(G)]
    lps = [c.large_lp(m, Cc, r, G) for r, c in enumerate(ctxs)]
    for lp in lps:
        lp.peer_attach_local(lps)
        lp.set_block(block)
        lp.set_window(window)
        lp.upload(*sf)
    outs = []
    for K in max_iters:
        st = [None] * G
        err = [None] * G

        def run(r):
            try:
                st[r] = lps[r].solve(K)
            except Exception as e:  # noqa: BLE001
                err[r] = e
        th = [threading.Thread(target=run, args=(r,)) for r in range(G)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not any(err), err
        assert len(set(st)) == 1, st
        parts = [lp.download(log_cap=1 << 16) for lp in lps]
        merged = dict(parts[0])
        merged["status"] = st[0]
        merged["tab"] = np.zeros((m, Cc))
        merged["tgtf"] = np.zeros(Cc)
        for lp, part in zip(lps, parts):
            sl = slice(lp.col0, lp.col0 + lp.local_cols)
            merged["tab"][:, sl] = part["tab"][:, sl]
            merged["tgtf"][sl] = part["tgtf"][sl]
            for k in ("eq2bv", "bv2eq", "nvset", "bvset", "log", "sol", "maxv"):  # replicated state
                assert np.array_equal(part[k], parts[0][k]), (k, lp.rank)
            assert part["iters"] == parts[0]["iters"]
        merged["checksums"] = [lp.checksum() for lp in lps]
        outs.append(merged)
    for lp in lps:
        lp.close()
    for c in ctxs:
        c.close()
    return outs


def assert_same_state(g, o, tag):
    assert g["status"] == o["status"], (tag, g["status"], o["status"])
    assert g["iters"] == o["iters"], (tag, g["iters"], o["iters"])
    assert np.array_equal(g["log"], o["log"][: len(g["log"])]), (tag, "pivot sequence")
    for k in ("eq2bv", "bv2eq", "nvset", "bvset"):
        assert np.array_equal(g[k], o[k]), (tag, k)
    for k in ("tab", "tgtf", "maxv", "sol"):
        assert np.array_equal(H.bits(g[k]), H.bits(o[k])), (tag, k)


def test_shard_bounds_cover_and_match_device():
    for Cc in (4, 5, 18, 19, 64, 1000, 16384):
        for G in (1, 2, 3, 4, 8):
            if (Cc + 1) // 2 < G:
                continue
            b = sharded.shard_bounds(Cc, G)
            assert b[0][0] == 0 and b[-1][1] == Cc
            assert all(b[r][1] == b[r + 1][0] for r in range(G - 1))
            assert all(lo % 2 == 0 and hi > lo for lo, hi in b)


@pytest.mark.parametrize("G", [2, 3, 4])
def test_inproc_dense_to_termination(G):
    seen = set()
    for seed, (m, n) in enumerate([(16, 15), (24, 23), (33, 20), (9, 9), (40, 64)]):
        leq, tg = H.gen_dense_lp(7000 + seed, m, n)
        sf = xp.slack_form(leq, tg)
        g = solve_sharded_inproc(G, sf, [H.NO_LIMIT], block=(0, 1, 5, 32)[seed % 4])[0]
        o = H.slack_solve_oracle("f64", *sf)
        assert_same_state(g, o, ("dense", G, m, n))
        seen.add(g["status"])
    assert seen <= {0, 1, 3}


@pytest.mark.parametrize("G", [2, 3, 4])
@pytest.mark.parametrize("window", [2, 1 << 20])
def test_inproc_windowed_leader(G, window):
    """Windowed panel on a sharded LP: rank 0 decides runs of pivots alone (its slice holds the
    window) and hands records + multiplier columns to the peers, which replay the replicated
    state; scans that leave the window go through the per-pivot exchange kernels."""
    for seed, (m, n) in enumerate([(16, 15), (24, 23), (33, 20), (40, 64)]):
        leq, tg = H.gen_dense_lp(7100 + seed, m, n)
        sf = xp.slack_form(leq, tg)
        g = solve_sharded_inproc(G, sf, [H.NO_LIMIT], block=(0, 32, 5, 1)[seed % 4], window=window)[0]
        assert_same_state(g, H.slack_solve_oracle("f64", *sf), ("wshard", G, window, m, n))
    leq, tg = H.gen_mixed_lp(5, 12, 30)
    leq[:, 30] = np.abs(leq[:, 30])
    sf = xp.slack_form(leq, tg)
    g = solve_sharded_inproc(G, sf, [H.NO_LIMIT], block=32, window=window)[0]
    assert_same_state(g, H.slack_solve_oracle("f64", *sf), ("wshard-mixed", G, window))
    leq, tg = H.gen_dense_lp(4243, 96, 95)
    sf = xp.slack_form(leq, tg)
    outs = solve_sharded_inproc(G, sf, [7, 40, 41], block=16, window=window)
    for K, g in zip((7, 40, 41), outs):
        assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=K), ("wshard-resume", G, window, K))


@pytest.mark.parametrize("G,m,n,window", [(2, 300, 1501, 512), (2, 257, 1300, 256), (3, 200, 2499, 512)])
def test_inproc_lookahead_leader(G, m, n, window):
    """Lookahead on the leader of a sharded LP: rank 0 closes a block at once and applies it to
    its slice beside the next k_wpanel (its slice is wider than the window), the peers keep the
    plain order.  Dense runs, a mixed-sign run (slow paths on the caught-up tableau), resumes."""
    for seed in range(2):
        leq, tg = H.gen_dense_lp(8700 + seed, m, n)
        sf = xp.slack_form(leq, tg)
        g = solve_sharded_inproc(G, sf, [400], block=(32, 0)[seed], window=window)[0]
        assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=400), ("lookshard", G, m, n, seed))
    leq, tg = H.gen_mixed_lp(8750, m, n)
    leq[:, n] = np.abs(leq[:, n])
    tg[:n] = np.abs(tg[:n])
    sf = xp.slack_form(leq, tg)
    outs = solve_sharded_inproc(G, sf, [45, 100, 333], block=32, window=window)
    for K, g in zip((45, 100, 333), outs):
        assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=K), ("lookshard-mixed", G, K))


@pytest.mark.parametrize("G", [1, 2, 3])
def test_upload_leq_on_shards(G):
    """xp_lp_f64_upload_leq on a (sharded) handle: every rank uploads only the columns of leq that
    fall into its slice and generates the slack columns on the device -- same state as uploading the
    host-built slack form."""
    for seed, (m, n) in enumerate([(16, 15), (33, 20), (12, 40)]):
        leq, tg = H.gen_dense_lp(7300 + seed, m, n)
        sf = xp.slack_form(leq, tg)
        ctxs = [xp.Context(0) for _ in range(G)]
        lps = [c.large_lp(m, n + m + 1, r, G) for r, c in enumerate(ctxs)]
        for lp in lps:
            if G > 1:
                lp.peer_attach_local(lps)
            lp.upload_leq(leq, tg)
        st = [None] * G
        th = [threading.Thread(target=lambda r=r: st.__setitem__(r, lps[r].solve(H.NO_LIMIT))) for r in range(G)]
        [t.start() for t in th]
        [t.join() for t in th]
        o = H.slack_solve_oracle("f64", *sf)
        for lp in lps:
            g = lp.download(log_cap=1 << 16)
            sl = slice(lp.col0, lp.col0 + lp.local_cols)
            assert st[lp.rank] == o["status"] and g["iters"] == o["iters"]
            assert np.array_equal(H.bits(g["tab"][:, sl]), H.bits(o["tab"][:, sl]))
            assert np.array_equal(H.bits(g["tgtf"][sl]), H.bits(o["tgtf"][sl]))
            assert np.array_equal(g["eq2bv"], o["eq2bv"]) and np.array_equal(H.bits(g["sol"]), H.bits(o["sol"]))
        for lp in lps:
            lp.close()
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("G", [2, 4])
def test_inproc_mixed_sign_slow_paths(G):
    """disableNV retries, the pass-2 ratio test and the findPivotNVandBVPair search
    all need columns that were not pre-extracted: the collective slow fetch."""
    seen = set()
    for seed in range(6):
        m, n = [(6, 5), (10, 9), (12, 30)][seed % 3]
        leq, tg = H.gen_mixed_lp(seed, m, n)
        leq[:, n] = np.abs(leq[:, n])
        sf = xp.slack_form(leq, tg)
        g = solve_sharded_inproc(G, sf, [H.NO_LIMIT], block=(0, 1, 7)[seed % 3])[0]
        assert_same_state(g, H.slack_solve_oracle("f64", *sf), ("mixed", G, seed))
        seen.add(g["status"])
    assert 1 in seen


def test_inproc_bounded_resume_and_checksum(ctx):
    """Run K pivots, resume to 3K: same bits as the oracle at each stop, and the
    shard checksums add up to the single-GPU checksum (mod 2^64)."""
    leq, tg = H.gen_dense_lp(4242, 96, 95)
    sf = xp.slack_form(leq, tg)
    outs = solve_sharded_inproc(4, sf, [5, 15], block=4)
    one = ctx.large_lp(*sf[0].shape)
    one.upload(*sf)
    for K, g in zip((5, 15), outs):
        assert_same_state(g, H.slack_solve_oracle("f64", *sf, max_iter=K), ("resume", K))
        one.solve(K)
        ref = one.checksum()
        for k in (0, 1):
            assert sum(c[k] for c in g["checksums"]) % (1 << 64) == ref[k]
    one.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multiprocess_ipc(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29650 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_WORKER_OK" in r.stdout


def test_multiprocess_ipc_one_gpu():
    """The CUDA-IPC path on a single-GPU box: two PROCESSES, both on cuda:0, exchange-block handles
    all-gathered between them and opened with cudaIpcOpenMemHandle -- the same code path as one
    process per GPU, minus NVLink.  Every rank compares its slice and the replicated state with the
    oracle bit for bit (full-width exchange kernels and the windowed leader)."""
    env = dict(os.environ, XP_TEST_ONE_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29649",
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_WORKER_OK" in r.stdout
