"""Dev tool: time the c2 batched kernel (device-resident inputs)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoly_b200 as xp
B, m, n = int(os.environ.get("B", 100000)), 32, 31
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(20261017)
leq = torch.rand((B, m, n + 1), dtype=torch.float64, device=dev, generator=g)
leq[:, :, n] = 1.0 + leq[:, :, n] * n
tg = torch.rand((B, n + 1), dtype=torch.float64, device=dev, generator=g); tg[:, n] = 0.0
status = torch.zeros(B, dtype=torch.int32, device=dev)
maxv = torch.zeros(B, dtype=torch.float64, device=dev)
piv = torch.zeros(B, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
ctx = xp.Context(0); lib = xp.lib()
def once():
    rc = lib.xp_six_two_stage_f64_batch_dev(ctx._h, B, m, n, C.c_void_p(leq.data_ptr()), C.c_void_p(tg.data_ptr()),
        C.c_uint32(xp.NO_ITER_LIMIT), 0, C.c_void_p(status.data_ptr()), C.c_void_p(maxv.data_ptr()), None, None, None, None,
        C.c_void_p(piv.data_ptr()))
    ctx.check(rc); return ctx.last_kernel_ms
for _ in range(2): once()
ms = float(np.median([once() for _ in range(int(os.environ.get("REPS", 5)))]))
tp = int(piv.cpu().numpy().astype(np.int64).sum())
print(f"{B / ms * 1e3:.0f} LPs/s  {ms:.2f} ms  pivots {tp}  {tp / ms * 1e3 / 1e6:.1f} Mpivots/s  max pivots/LP {int(piv.max())}")
