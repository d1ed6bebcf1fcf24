"""Dev probe: mixed-sign small LP, windowed, to termination; a watchdog thread dumps the device state."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H
import xpoly_b200 as xp

m, n, window, block, seed = 6, 5, int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
leq, tg = H.gen_mixed_lp(seed, m, n)
leq[:, n] = np.abs(leq[:, n])
sf = xp.slack_form(leq, tg)
o = H.slack_solve_oracle("f64", *sf)
print("oracle status", o["status"], "iters", o["iters"], flush=True)
ctx = xp.Context(0)
lp = ctx.large_lp(*sf[0].shape)
lp.set_window(window)
lp.set_block(block)
lp.upload(*sf)
lib = xp.lib()
names = "status cnt t kblk blk q slow pivot_pending wseq wb_pending rest_pending rest_slot n_touched wcnt qmax wfail".split()
done = []


def watch():
    for k in range(6):
        time.sleep(1.0)
        if done:
            return
        out = (C.c_longlong * 16)()
        lib.xp_lp_f64_debug_state(lp._h, out)
        print("watch", k, dict(zip(names, list(out))), "launches", ctx.launches, flush=True)
    os._exit(3)


threading.Thread(target=watch, daemon=True).start()
st = lp.solve(H.NO_LIMIT)
done.append(1)
print("status", st, "launches", ctx.launches)
