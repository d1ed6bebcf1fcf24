"""Dev probe: in-process sharded LP with the lookahead on the leader; dumps the device state on failure."""
import ctypes as C
import os
import sys
import threading

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H
import xpoly_b200 as xp

G, m, n, window = 2, 300, 1501, 512
kind, seed, block, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
if kind == "dense":
    leq, tg = H.gen_dense_lp(seed, m, n)
else:
    leq, tg = H.gen_mixed_lp(seed, m, n)
    leq[:, n] = np.abs(leq[:, n])
    tg[:n] = np.abs(tg[:n])
sf = xp.slack_form(leq, tg)
ctxs = [xp.Context(0) for _ in range(G)]
lps = [c.large_lp(m, sf[0].shape[1], r, G) for r, c in enumerate(ctxs)]
for lp in lps:
    lp.peer_attach_local(lps)
    lp.set_block(block)
    lp.set_window(window)
    lp.upload(*sf)
st = [None] * G


def run(r):
    try:
        st[r] = lps[r].solve(K)
    except Exception as e:  # noqa: BLE001
        st[r] = repr(e)


th = [threading.Thread(target=run, args=(r,)) for r in range(G)]
[t.start() for t in th]
[t.join() for t in th]
print("status", st)
lib = xp.lib()
names = "status cnt t kblk blk q slow pivot_pending wseq wb_pending rest_pending rest_slot n_touched wcnt qmax wfail".split()
for r, lp in enumerate(lps):
    out = (C.c_longlong * 16)()
    lib.xp_lp_f64_debug_state(lp._h, out)
    print(r, dict(zip(names, list(out))))
