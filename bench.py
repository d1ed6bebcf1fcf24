#!/usr/bin/env python
"""bench.py -- pivots/s of the simplex hot path on the config-3 workload
(dense FP64 LP, 8192 x 16384 tableau) plus the config-2 batched figure.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K ...  (reference CPU arm)

A "step" is one pass of the hot path over one batch of work: P simplex
iterations (pricing -> ratio test -> rank-1 pivot) of the HBM-resident tableau.
`value` is device-resident (inputs in HBM when the timed region starts); `e2e`
goes through the C-ABI call that stands where the reference arm's TwoStageMethod(leq, tgtf)
stands (xp_six_two_stage_f64_large) with pinned HOST buffers, host<->device copies inside the
timed region (`e2e_slack`: the kernel-level xp_six_slack_f64, whole tableau up and down).
Two JSON lines are printed: the secondary legs first (c2 / c4 / c5 / has_solution), the
main line LAST.  Inputs (1 GiB) are
larger than L2 (126 MB), so no flush is needed between iterations.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_ROWS = 8192          # config 3: leq 8192 x 8192 (8191 vars + rhs) -> tableau 8192 x 16384
N_VARS = 8191
BATCH_LPS = 100_000    # config 2: 100k LPs, leq 32 x 32 (31 vars) -> tableau 32 x 64
BATCH_M, BATCH_N = 32, 31
SEED = 20261017


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region.  The sampler is started
    early (nvidia-smi needs a second or two before its first line), every line is stamped on
    arrival, and `window(t0, t1)` summarises the lines that arrived inside the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "10"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_ready(self, timeout=10.0):
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def window(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0,
                    "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for ts, ln in list(self.lines):
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- CPU side
def cpu_pivot_rate(leq, tgtf, pivots, prefer_ref, setup_s=None):
    """Pivots/s of the reference's CPU solver on the same LP, 1 thread (the
    reference has no intra-LP parallelism).  kind 'reference' = oracle/_ref (the
    unmodified reference; solve loop timed inside the checker; reference set-up measured with max_iter=0 and subtracted),
    kind 'port' = oracle/xp_oracle.c solveSlackForm restatement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    import xpoly_b200 as xp
    if prefer_ref and H.ref() is not None:
        # SIX::TwoStageMethod timed inside the shim (excludes matrix fill / dump);
        # its slack-form set-up is measured with max_iter=0 and subtracted.
        H.ref().ref_last_two_stage_seconds.restype = C.c_double
        t_setup = setup_s
        if t_setup is None:  # warm minimum of two set-up-only runs
            ts = []
            for _ in range(2):
                H.two_stage("ref", "f64", leq, tgtf, 0)
                ts.append(H.ref().ref_last_two_stage_seconds())
            t_setup = min(ts)
        r = H.two_stage("ref", "f64", leq, tgtf, pivots)
        t_run = H.ref().ref_last_two_stage_seconds()
        dt = max(t_run - t_setup, 1e-9)
        return pivots / dt, "reference", dict(eq2bv=r["eq2bv"], setup_s=t_setup, run_s=t_run)
    sf = xp.slack_form(leq, tgtf)
    H.oracle().xo_last_solve_seconds.restype = C.c_double
    r = H.slack_solve_oracle("f64", *sf, max_iter=pivots, log_cap=0)
    dt = max(H.oracle().xo_last_solve_seconds(), 1e-9)  # the solve loop alone
    return pivots / dt, "port", dict(eq2bv=r["eq2bv"], setup_s=0.0, run_s=dt)


def run_reference_arm(args):
    """The reference's own CPU solver (oracle/_ref, else the oracle port) on the c3 LP, one host
    thread (the reference has no intra-LP parallelism).  A step is a bounded sample of the
    workload: `--ref-pivots` simplex iterations.  One call of the reference's TwoStageMethod on
    this LP costs ~12 s of set-up (it materialises an (n+m)^2 `vc` matrix, lpsol.h:1415), so the
    W warm-up and K timed steps are run as two calls of one deterministic pivot sequence --
    max_iter = P*W and max_iter = P*(W+K) -- and the timed region is their difference."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from xpoly_b200.synth import dense_lp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    import xpoly_b200 as xp
    leq, tgtf = dense_lp(SEED, args.m, args.n)
    P, W, K = args.ref_pivots, args.warmup, args.steps
    use_ref = (not args.port) and H.ref() is not None
    if use_ref:
        H.ref().ref_last_two_stage_seconds.restype = C.c_double

        def run(pivots):  # seconds inside SIX::TwoStageMethod (set-up + `pivots` iterations)
            H.two_stage("ref", "f64", leq, tgtf, pivots)
            return H.ref().ref_last_two_stage_seconds()
        kind = "reference"
    else:
        sf = xp.slack_form(leq, tgtf)
        H.oracle().xo_last_solve_seconds.restype = C.c_double

        def run(pivots):  # seconds inside the port's solveSlackForm loop
            H.slack_solve_oracle("f64", *sf, max_iter=pivots, log_cap=0)
            return H.oracle().xo_last_solve_seconds()
        kind = "port"
    t_w = run(P * W)
    t_all = run(P * (W + K))
    dt = max(t_all - t_w, 1e-9)
    value = P * K / dt
    C_cols = args.n + args.m + 1
    line = {
        "impl": "reference", "metric": "pivots/s", "value": value, "unit": "pivots/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1000.0 * dt / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"c3: dense FP64 LP, tableau {args.m}x{C_cols}, simplex iterations "
                               "under the reference pivot rule from the slack basis",
                   "pivots_per_step": P,
                   "sample": f"bounded: {P} pivots per step (the reference needs ~0.25-0.8 s per pivot here)"},
        "cpu_baseline": {"value": value, "unit": "pivots/s", "cores": 1, "kind": kind,
                         "sample": f"{P} pivots of the same {args.m}x{C_cols} LP per step; timed region = "
                                   f"TwoStageMethod(max_iter={P * (W + K)}) - TwoStageMethod(max_iter={P * W}), "
                                   "both timed inside the checker (same deterministic pivot sequence)"},
        "e2e": {"value": value, "unit": "pivots/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------- GPU side
def ctx_sm_count(torch, dev):
    return torch.cuda.get_device_properties(dev).multi_processor_count


def sm_clock_hz(torch, dev):
    """Max SM clock (MEASURED_PEAKS.json if present, else the device property)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["sm_max_mhz"]) * 1e6
    except Exception:
        return torch.cuda.get_device_properties(dev).clock_rate * 1e3


def run_batched(ctx, xp, torch, dev, with_cpu=True, rank=0, world=1, dist=None):
    """Config 2: 100k LPs of tableau 32x64, one warp per LP.  With N ranks the batch is split
    N ways (independent units, no collective: SURVEY 8e); times are the max over ranks."""
    Btot, m, n = BATCH_LPS, BATCH_M, BATCH_N
    B = Btot * (rank + 1) // world - Btot * rank // world

    def sync_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    # LP k of the batch comes from std::mt19937_64(2024 + k) (A_ij ~ U(0,1) row by row
    # with b_i = 1 + U * n, then c_j ~ U(0,1)); generated on the host, outside any timed region
    from xpoly_b200 import synth
    lo_k = Btot * rank // world
    h_leq, h_tg = synth.dense_lp_batch(2024 + lo_k, B, m, n)
    leq = torch.from_numpy(h_leq).to(dev)
    tg = torch.from_numpy(h_tg).to(dev)
    del h_leq, h_tg
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    maxv = torch.zeros(B, dtype=torch.float64, device=dev)
    pivots = torch.zeros(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    lib = xp.lib()

    def once():
        rc = lib.xp_six_two_stage_f64_batch_dev(
            ctx._h, B, m, n, C.c_void_p(leq.data_ptr()), C.c_void_p(tg.data_ptr()),
            C.c_uint32(xp.NO_ITER_LIMIT), 0, C.c_void_p(status.data_ptr()),
            C.c_void_p(maxv.data_ptr()), None, None, None, None, C.c_void_p(pivots.data_ptr()))
        ctx.check(rc)
        return ctx.last_kernel_ms
    for _ in range(3):
        once()
    ms = []
    for _ in range(5):
        barrier()
        ms.append(sync_max(once()))
    dev_ms = float(np.median(ms))
    st = status.cpu().numpy()
    tot_piv = int(sync_sum(float(pivots.cpu().numpy().astype(np.int64).sum())))
    # end to end through the host-pointer C-ABI call (H2D of all LPs from PINNED host memory,
    # D2H of results)
    def pinned_like(t):
        hp = C.c_void_p()
        ctx.check(lib.xp_host_alloc(ctx._h, C.c_size_t(t.numel() * 8), C.byref(hp)))
        a = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=tuple(t.shape))
        a[...] = t.cpu().numpy()
        return a, hp
    h_leq, hp1 = pinned_like(leq)
    h_tg, hp2 = pinned_like(tg)
    t_e2e = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        out = ctx.two_stage_f64_batch(h_leq, h_tg, want=("status", "maxv"))
        t_e2e.append(sync_max(time.perf_counter() - t0))
    e2e_s = float(np.median(t_e2e))
    e2e_status = out["status"].copy()
    h2d_bytes = int(sync_sum(float(h_leq.nbytes + h_tg.nbytes)))
    cpu = None
    if with_cpu and world == 1:
        # CPU baseline beside it: the oracle port's TwoStageMethod on a bounded
        # sample of the same LPs, one LP after the other per thread, on 1 core and on all cores.
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import harness as H
        o = H.oracle()
        o.xo_two_stage_f64_many.restype = C.c_double
        T = max(1, min(64, len(os.sched_getaffinity(0))))
        S1, ST = min(B, 4000), min(B, 4000 * T)
        cst = np.zeros(ST, dtype=np.int32)
        cmx = np.zeros(ST)

        def cpu_run(lo, hi):
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            return o.xo_two_stage_f64_many(hi - lo, m, n, vp(h_leq[lo:hi]), vp(h_tg[lo:hi]),
                                           vp(cst[lo:hi]), vp(cmx[lo:hi]))
        t1 = cpu_run(0, S1)
        t0 = time.perf_counter()
        th = [threading.Thread(target=cpu_run, args=(ST * i // T, ST * (i + 1) // T)) for i in range(T)]
        [t.start() for t in th]
        [t.join() for t in th]
        tT = time.perf_counter() - t0
        cpu = {"value": ST / tT, "unit": "LPs/s", "cores": T, "kind": "port",
               "sample": f"the first {ST} of the {B} LPs, {T} host threads each looping over its "
                         "slice (the reference solves one LP per call)",
               "one_core": {"value": S1 / t1, "sample": f"the first {S1} LPs"},
               "status_matches_gpu": bool(np.array_equal(cst, e2e_status[:ST])),
               "objective_bits_match_gpu": bool(np.array_equal(cmx.view(np.uint64),
                                                               out["maxv"][:ST].view(np.uint64)))}
    del h_leq, h_tg
    ctx.check(lib.xp_host_free(ctx._h, hp1))
    ctx.check(lib.xp_host_free(ctx._h, hp2))
    # roofline of this path: the FP64 pipe.  One pivot is (m+1) x C non-fused multiply + add
    # pairs (lpsol.h:1481-1501); the pipe issues 64 FP64 lanes per clock per SM.
    fp64_ops = tot_piv * 2.0 * (m + 1) * (n + m + 1)
    fp64_peak = ctx_sm_count(torch, dev) * 64 * sm_clock_hz(torch, dev)
    return {
        "metric": "small LPs/s", "workload": f"c2: {Btot} LPs, tableau {m}x{n + m + 1}, FP64, LP k drawn from std::mt19937_64(2024 + k)"
                                             + (f", split over {world} GPUs" if world > 1 else ""),
        "value": Btot / (dev_ms * 1e-3), "unit": "LPs/s", "ms": dev_ms, "scaling": "strong",
        "pivots_total": tot_piv, "pivots_per_s": tot_piv / (dev_ms * 1e-3),
        "kernel": "k_warp_f64<32,2>: one warp per LP, tableau in registers",
        "roofline": {"bound": "fp64 pipe (non-fused mul + add, 64 lanes/clk/SM)",
                     "achieved": fp64_ops / (dev_ms * 1e-3) / 1e12, "peak": world * fp64_peak / 1e12,
                     "unit": "Tops/s", "frac": fp64_ops / (dev_ms * 1e-3) / (world * fp64_peak),
                     "algorithmic_ops_per_pivot": 2 * (m + 1) * (n + m + 1)},
        "status_mix_rank0": {str(k): int((st == k).sum()) for k in np.unique(st)},
        "e2e": {"value": Btot / e2e_s, "unit": "LPs/s",
                "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(Btot * 12),
                "api": "xp_six_two_stage_f64_batch (pinned host buffers; chunks uploaded while "
                       "the previous chunk is being solved)"},
        "status_matches_e2e": bool(np.array_equal(e2e_status, st)),
        "cpu_baseline": cpu,
    }


def _threads(fn, N, T):
    """fn(lo, hi) on T host threads over [0, N) (the oracle's C loops release the GIL); wall seconds."""
    t0 = time.perf_counter()
    th = [threading.Thread(target=fn, args=(N * i // T, N * (i + 1) // T)) for i in range(T)]
    [t.start() for t in th]
    [t.join() for t in th]
    return time.perf_counter() - t0


def run_exact_and_bnb(ctx, xp, torch, dev, rank=0, world=1, dist=None, with_cpu=True):
    """Config 4 (10k exact 24x48 LPs, fraction-free int64), config 5 (knapsack-style B&B, node
    relaxations batched on the GPU; the reference only solves this family up to ~50 variables,
    SURVEY 8d) and Lineq::has_solution batches, through the host-pointer C ABI.  With N ranks the
    independent LPs / trees / systems are split N ways (no collective); times are the max
    over ranks.  CPU baselines (oracle port, 1 core and all cores) and an in-run comparison of a
    sample with the oracle ride along at N = 1."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    out = {}
    T = max(1, min(64, len(os.sched_getaffinity(0))))

    def mx(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sm(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def part(n):
        return n * rank // world, n * (rank + 1) // world
    r = np.random.RandomState(777)
    B, m, n = 10_000, 24, 23
    A = r.randint(0, 4, size=(B, m, n)) * (r.uniform(size=(B, m, n)) < 0.3)
    leq = np.zeros((B, m, n + 1), dtype=np.int64)
    leq[:, :, :n] = A
    leq[:, :, n] = r.randint(0, 21, size=(B, m))
    tg = np.zeros((B, n + 1), dtype=np.int64)
    tg[:, :n] = r.randint(1, 6, size=(B, n))
    lo, hi = part(B)
    ctx.two_stage_i64_batch(leq[:64], tg[:64])
    ts, dms = [], []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        res = ctx.two_stage_i64_batch(leq[lo:hi], tg[lo:hi])
        ts.append(mx(time.perf_counter() - t0))
        dms.append(mx(ctx.last_kernel_ms))
    st = res["status"]
    piv = int(sm(res["pivots"].astype(np.int64).sum()))
    dev_ms = float(np.median(dms))
    smem_peak = world * ctx_sm_count(torch, dev) * 128 * sm_clock_hz(torch, dev)  # 128 B/clk/SM
    smem_bytes = piv * 2.0 * (m + 1) * (n + m + 1) * 8
    out["exact"] = {"metric": "exact LPs/s", "workload": f"c4: {B} LPs, tableau {m}x{n + m + 1}, fraction-free "
                    "int64 entries / 128-bit products, e2e through host pointers"
                    + (f", split over {world} GPUs" if world > 1 else ""),
                    "value": B / float(np.median(ts)), "unit": "LPs/s", "device_ms": dev_ms,
                    "device_LPs_per_s": B / (dev_ms * 1e-3), "pivots_total": piv,
                    "roofline": {"bound": "shared-memory bandwidth (tableau resident in one CTA's shared memory: "
                                          "2 (m+1) C 8 B per pivot)", "achieved": smem_bytes / (dev_ms * 1e-3) / 1e9,
                                 "peak": smem_peak / 1e9, "unit": "GB/s",
                                 "frac": smem_bytes / (dev_ms * 1e-3) / smem_peak,
                                 "note": "latency-bound: the batch ends with its longest LP on one CTA"},
                    "status_mix_rank0": {str(k): int((st == k).sum()) for k in np.unique(st)}}
    if with_cpu and world == 1:
        o = H.oracle()
        o.xo_two_stage_rat_many.restype = C.c_double
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        S1, ST = 400, min(B, 400 * T)
        rl, rt = H.to_rat(leq[:ST]), H.to_rat(tg[:ST])
        cst, cmv, cap = np.zeros(ST, dtype=np.int32), np.zeros((ST, 2), dtype=np.int32), np.zeros(ST, dtype=np.uint8)
        t1 = o.xo_two_stage_rat_many(S1, m, n, vp(rl), vp(rt), vp(cst), vp(cmv), vp(cap))
        # single-threaded pass: per LP, did the reference's lossy appro() fire?  (those are not comparable)
        ok = cap[:S1] == 0
        same = bool(np.array_equal(cst[:S1][ok], st[:S1][ok]))
        for k in np.nonzero(ok & (cst[:S1] == 0))[0]:
            same &= int(res["maxv"][k][0]) * int(cmv[k][1]) == int(cmv[k][0]) * int(res["maxv"][k][1])
        tT = _threads(lambda a, b: o.xo_two_stage_rat_many(b - a, m, n, vp(rl[a:b]), vp(rt[a:b]), vp(cst[a:b]),
                                                           vp(cmv[a:b]), None), ST, T)
        out["exact"]["cpu_baseline"] = {"value": ST / tT, "unit": "LPs/s", "cores": T, "kind": "port",
                                        "sample": f"the first {ST} LPs, {T} host threads each looping over its slice",
                                        "one_core": {"value": S1 / t1, "sample": f"the first {S1} LPs"}}
        out["exact"]["matches_oracle_sample"] = same
        out["exact"]["sample_compared"] = int(ok.sum())
        out["exact"]["sample_skipped_reference_inexact"] = int((~ok).sum())
    # c5: 256 independent 40-item knapsacks (1 + n rows), general-integer B&B
    Tn, nk = 256, 40
    w = r.randint(5, 41, size=(Tn, nk))
    pr = r.randint(5, 61, size=(Tn, nk))
    L = np.zeros((Tn, nk + 1, nk + 1), dtype=np.int64)
    L[:, 0, :nk] = w
    L[:, 0, nk] = w.sum(axis=1) // 3
    for j in range(nk):
        L[:, 1 + j, j] = 1
        L[:, 1 + j, nk] = 1
    Gm = np.zeros((Tn, nk + 1), dtype=np.int64)
    Gm[:, :nk] = pr
    lo, hi = part(Tn)
    ctx.mip_solve_rat_batch(0, 0, L[:4], Gm[:4])
    barrier()
    t0 = time.perf_counter()
    mres = ctx.mip_solve_rat_batch(0, 0, L[lo:hi], Gm[lo:hi])
    dt = mx(time.perf_counter() - t0)
    nodes = int(sm(mres["nodes"].astype(np.int64).sum()))
    ms_ = mres["status"]
    out["bnb"] = {"metric": "B&B node LPs/s", "workload": f"c5: {Tn} knapsack MIPs of {nk} items "
                  f"(tableau {nk + 1}x{2 * nk + 2} at the root), trees advanced in lockstep, node "
                  "relaxations batched on the GPU, decisions replayed in the reference's DFS order"
                  + (f", trees split over {world} GPUs" if world > 1 else ""),
                  "value": nodes / dt, "unit": "node LPs/s", "trees_per_s": Tn / dt, "nodes_total": nodes,
                  "status_mix_rank0": {str(k): int((ms_ == k).sum()) for k in np.unique(ms_)}}
    if with_cpu and world == 1:
        o.xo_mip_solve_rat_many.restype = C.c_double
        S1, ST = 8, min(Tn, 4 * T)
        rl, rt = H.to_rat(L[:ST]), H.to_rat(Gm[:ST])
        cst, cv, cn = np.zeros(ST, dtype=np.int32), np.zeros((ST, 2), dtype=np.int32), np.zeros(ST, dtype=np.int32)
        t1 = o.xo_mip_solve_rat_many(S1, 0, 0, nk + 1, nk, vp(rl), vp(rt), vp(cst), vp(cv), vp(cn))
        same = bool(np.array_equal(cst[:S1], ms_[:S1]) and np.array_equal(cn[:S1], mres["nodes"][:S1]))
        for k in range(S1):
            if cst[k] == 0:
                same &= int(cv[k][0]) * int(mres["v"][k][1]) == int(mres["v"][k][0]) * int(cv[k][1])
        n1 = int(cn[:S1].sum())
        tT = _threads(lambda a, b: o.xo_mip_solve_rat_many(b - a, 0, 0, nk + 1, nk, vp(rl[a:b]), vp(rt[a:b]),
                                                           vp(cst[a:b]), vp(cv[a:b]), vp(cn[a:b])), ST, T)
        out["bnb"]["cpu_baseline"] = {"value": int(cn.sum()) / tT, "unit": "node LPs/s", "cores": T, "kind": "port",
                                      "sample": f"the first {ST} trees, {T} host threads each looping over its slice",
                                      "one_core": {"value": n1 / t1, "sample": f"the first {S1} trees"}}
        out["bnb"]["matches_oracle_sample"] = same
    # 8(f1/f2): Lineq::has_solution over a dependence-graph build's worth of queries (systems of
    # different sizes, integer solutions wanted), one ragged call; the oracle port beside it
    Q = 20_000
    systems = []
    for k in range(Q):
        nq, mq = int(r.randint(2, 6)), int(r.randint(3, 10))
        sysm = np.zeros((mq, nq + 1), dtype=np.int64)
        sysm[:, :nq] = r.randint(-2, 4, size=(mq, nq)) * (r.uniform(size=(mq, nq)) < 0.7)
        sysm[:, nq] = r.randint(0, 25, size=mq)
        systems.append((sysm, None))
    lo, hi = part(Q)
    # the C-ABI arguments of xp_has_solution_rat_ragged, packed outside the timed region (a C++
    # producer appends rows to such pools as it goes; packing 20 000 numpy arrays in Python is not
    # part of the path)
    hs_m = np.array([sy[0].shape[0] for sy in systems[lo:hi]], dtype=np.int32)
    hs_n = np.array([sy[0].shape[1] - 1 for sy in systems[lo:hi]], dtype=np.int32)
    hs_off = np.zeros(hi - lo, dtype=np.int64)
    hs_off[1:] = np.cumsum(hs_m[:-1].astype(np.int64) * (hs_n[:-1] + 1))
    hs_pool = np.ascontiguousarray(np.concatenate([H.to_rat(sy[0]).reshape(-1, 2) for sy in systems[lo:hi]]))
    res = np.zeros(hi - lo, dtype=np.int32)
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib = xp.lib()

    def hs_call():
        ctx.check(lib.xp_has_solution_rat_ragged(ctx._h, hi - lo, hp(hs_n), hp(hs_m), hp(hs_off), hp(hs_pool),
                                                 C.c_size_t(len(hs_pool)), None, None, None, C.c_size_t(0), 1, 1, hp(res)))
    hs_call()
    ts = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        hs_call()
        ts.append(mx(time.perf_counter() - t0))
    dt = float(np.median(ts))
    hs_dev_ms = mx(ctx.last_kernel_ms)
    out["has_solution"] = {"metric": "dependence queries/s", "workload": f"{Q} Lineq::has_solution systems, "
                           "2-5 variables x 3-9 inequalities, integer solutions (max then min MIP with branch & bound), "
                           "one xp_has_solution_rat_ragged call from host memory"
                           + (f" per rank, split over {world} GPUs" if world > 1 else ""),
                           "value": Q / dt, "unit": "queries/s", "feasible": int(sm((res == 1).sum())),
                           "device_ms": hs_dev_ms, "device_queries_per_s": Q / (hs_dev_ms * 1e-3) if hs_dev_ms > 0 else None,
                           "kernel": "k_has_solution: one warp per query, both MIPs incl. branch & bound on the device"}
    if with_cpu and world == 1:
        o.xo_has_solution_rat_many.restype = C.c_double
        S1, ST = 1000, min(Q, 1000 * T)
        msq = np.array([sy[0].shape[0] for sy in systems[:ST]], dtype=np.int32)
        nsq = np.array([sy[0].shape[1] - 1 for sy in systems[:ST]], dtype=np.int32)
        off = np.zeros(ST, dtype=np.int64)
        off[1:] = np.cumsum(msq[:-1].astype(np.int64) * (nsq[:-1] + 1))
        pool = np.ascontiguousarray(np.concatenate([H.to_rat(sy[0]).reshape(-1, 2) for sy in systems[:ST]]))
        cres = np.zeros(ST, dtype=np.int32)
        t1 = o.xo_has_solution_rat_many(S1, vp(msq), vp(nsq), vp(off), vp(pool), 1, 1, vp(cres))
        same = bool(np.array_equal(cres[:S1], res[:S1]))
        tT = _threads(lambda a, b: o.xo_has_solution_rat_many(b - a, vp(msq[a:b]), vp(nsq[a:b]), vp(off[a:b]), vp(pool),
                                                              1, 1, vp(cres[a:b])), ST, T)
        out["has_solution"]["cpu_baseline"] = {"value": ST / tT, "unit": "queries/s", "cores": T, "kind": "port",
                                               "sample": f"the first {ST} systems, {T} host threads each looping over its slice",
                                               "one_core": {"value": S1 / t1, "sample": f"the first {S1} systems"}}
        out["has_solution"]["matches_oracle_sample"] = same
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import xpoly_b200 as xp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (xpoly_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_sum_u64(x):
        """Sum of one uint64 per rank, mod 2^64."""
        if world == 1:
            return int(x) % (1 << 64)
        t = torch.tensor([int(x) & 0xFFFFFFFF, int(x) >> 32], dtype=torch.int64, device=dev)
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        return sum(int(v[0].item()) + (int(v[1].item()) << 32) for v in g) % (1 << 64)

    def bcast_flag(ok):
        """rank 0's verdict to everyone (so that every rank exits the same way)."""
        if world == 1:
            return ok
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.broadcast(t, src=0)
        return bool(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx = xp.Context(local_rank)
    lib = xp.lib()
    m, n = args.m, args.n
    Ccols = n + m + 1
    P = args.pivots
    # the binding roof of the tableau pass at k >= ~23: separate DMUL + DADD (no DFMA: the reference
    # rounds twice), measured on this device now
    tops, pms = C.c_double(0), C.c_double(0)
    ctx.check(lib.xp_probe_fp64_nonfused(ctx._h, C.byref(tops), C.byref(pms)))
    fp64_peak = tops.value * 1e12
    fp64_nominal = ctx_sm_count(torch, dev) * 64 * sm_clock_hz(torch, dev)
    if world > 1:
        from xpoly_b200 import sharded
        lp = sharded.ShardedLP(ctx, m, Ccols, rank, world, dist)
    else:
        lp = ctx.large_lp(m, Ccols)
    if args.window is not None:
        lp.set_window(args.window)
    local_cols = lp.local_cols
    peak, peak_src = measured_peak_gbs()
    B_pivot = 2.0 * (m + 1) * local_cols * 8  # read+write of every entry, per pivot

    def timed_run(block, pivots, steps, warmup, load_s=0.0):
        """`steps` steps of `pivots` simplex iterations each; device time = max over ranks."""
        lp.set_block(block)
        lp.fill_synthetic(SEED)
        done = 0

        def step():
            nonlocal done
            st = lp.solve(done + pivots)
            done += pivots
            if st != xp.SIX_TIME_OUT:  # LP terminated: restart from a fresh instance
                lp.fill_synthetic(SEED + done)
                done = 0
            return ctx.last_kernel_ms
        for _ in range(warmup):
            step()
        lib.xp_lp_f64_profile(lp._h, 1)
        barrier()
        l0 = ctx.launches
        t0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(steps):
            dev_ms += step()
        barrier()
        t1 = time.perf_counter()
        wall = t1 - t0
        launches = ctx.launches - l0
        if world > 1:
            t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms = float(t.item())
        n_sw, sw_ms, gap_ms = C.c_uint64(0), C.c_double(0), C.c_double(0)
        lib.xp_lp_f64_profile_read(lp._h, C.byref(n_sw), C.byref(sw_ms), C.byref(gap_ms))
        lib.xp_lp_f64_profile(lp._h, 0)
        # The timed region can be shorter than nvidia-smi's sampling period (8 GPUs: a few ms
        # per step): keep the identical load running, untimed, until the clock sampler has had
        # `load_s` seconds of it.  Every rank derives the same count from the reduced time.
        extra = 0
        if load_s > 0 and dev_ms > 0 and dev_ms * 1e-3 < load_s:
            extra = int(np.ceil((load_s - dev_ms * 1e-3) / (dev_ms * 1e-3 / steps)))
            for _ in range(extra):
                step()
            barrier()
        t2 = time.perf_counter()
        return dict(dev_ms=dev_ms, wall=wall, launches=launches, pivots=steps * pivots,
                    flushes=int(n_sw.value), flush_ms=sw_ms.value, gap_ms=gap_ms.value,
                    t0=t0, t1=t1, t2=t2, extra_steps=extra)

    sampler.wait_ready()
    main = timed_run(args.block, P, args.steps, args.warmup, load_s=0.5)
    shared_sms = int(lib.xp_lp_f64_pass_shared_sms(lp._h))  # (sharded: rank 0 -- the rank whose launches are timed here)
    sampler.stop()
    clocks = sampler.window(main["t0"], main["t2"])
    clocks["window"] = ("timed region" if main["extra_steps"] == 0 else
                        f"timed region + {main['extra_steps']} untimed steps of the identical load")
    value = main["pivots"] / (main["dev_ms"] * 1e-3)
    k_eff = main["pivots"] / max(main["flushes"], 1)  # pivots applied per k_flush launch
    flush_avg_ms = main["flush_ms"] / max(main["flushes"], 1)
    moved = 2.0 * m * local_cols * 8  # one read + one write of the slice per launch
    traffic = None
    # (the kernel as the timed schedule launches it: replaying out of the ring beside the cluster
    # when the lookahead is on, the plain pass otherwise)
    for prof in (("r02_flush_lag_ncu_full.json",) if shared_sms else ()) + ("r02_flush_ncu_full.json", "r01_flush_ncu_full.json"):
        pth = os.path.join(ROOT, "profiles", prof)
        if os.path.exists(pth) and world == 1:
            try:
                traffic = json.load(open(pth)).get("dram_bytes_per_launch")
                break
            except Exception:
                traffic = None
    # the same blocked run with the lookahead off: every pass has the device to itself (the kernel's
    # own roofline figure, beside the one of the schedule `value` is measured on)
    alone = None
    shared_any = shared_sms  # (sharded: only the leader reports it, every rank must run the same steps)
    if world > 1:
        t = torch.tensor([shared_sms], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shared_any = int(t.item())
    if shared_any:
        os.environ["XP_NO_LOOKAHEAD"] = "1"
        try:
            alone = timed_run(args.block, P, 5, 3)
        finally:
            del os.environ["XP_NO_LOOKAHEAD"]
    # the reference's own schedule (one tableau pass per pivot), same kernels with k = 1
    r1 = timed_run(1, args.rank1_pivots, 3, 1)
    r1_value = r1["pivots"] / (r1["dev_ms"] * 1e-3)
    r1_sweep_ms = r1["flush_ms"] / max(r1["flushes"], 1)
    fp64_ops = k_eff * 2.0 * m * local_cols  # non-fused operations of one launch (DMUL + DADD per entry and pivot)
    ach = fp64_ops / (flush_avg_ms * 1e-3) if flush_avg_ms > 0 else 0.0
    n_sm = ctx_sm_count(torch, dev)
    roofline = {
        "bound": "fp64_pipe_nonfused", "kernel": "k_flush_w (rank-k tableau pass, k pivots per launch, DMUL + DADD per entry and pivot)"
                 + (f"; lookahead: each pass runs beside the {shared_sms}-CTA k_wpanel cluster deciding the next block and is timed "
                    "from the cluster's launch to its own end" if shared_sms else ""),
        "achieved": ach / 1e12, "peak": fp64_peak / 1e12, "unit": "Tops/s",
        "frac": ach / fp64_peak if fp64_peak > 0 else None,
        "sms_shared_with_panel": shared_sms,
        "frac_of_sms_held": (ach / (fp64_peak * (n_sm - shared_sms) / n_sm)) if fp64_peak > 0 else None,
        "peak_source": f"measured in this run (xp_probe_fp64_nonfused, {pms.value:.1f} ms; nominal 64 lanes/clk/SM = "
                       f"{fp64_nominal / 1e12:.2f})",
        "traffic": traffic, "ops_per_launch": fp64_ops, "pivots_per_launch": k_eff,
        "avg_launch_ms": flush_avg_ms, "launches_timed": main["flushes"],
        "flush_share_of_step": main["flush_ms"] / main["dev_ms"] if main["dev_ms"] > 0 else None,
        "step_frac": value * 2.0 * m * local_cols / fp64_peak if fp64_peak > 0 else None,
        "hbm_peak_GBps": peak, "hbm_peak_source": peak_src,
        "hbm_moved_frac": (moved / (flush_avg_ms * 1e-3) / 1e9 / peak) if flush_avg_ms > 0 else None,
        "algorithmic_hbm_equiv_frac": (k_eff * B_pivot / (flush_avg_ms * 1e-3) / 1e9 / peak) if flush_avg_ms > 0 else None,
        "rank1_pivots_per_s": r1_value,
        "rank1_kernel_hbm_frac": (B_pivot / (r1_sweep_ms * 1e-3) / 1e9 / peak) if r1_sweep_ms > 0 else None,
        "rank1_whole_pivot_frac_of_8TBps": r1_value * B_pivot / 8e12}
    if alone and alone["flushes"] and alone["flush_ms"] > 0:
        a_ops = alone["pivots"] / alone["flushes"] * 2.0 * m * local_cols
        a_ms = alone["flush_ms"] / alone["flushes"]
        roofline["alone_frac"] = a_ops / (a_ms * 1e-3) / fp64_peak if fp64_peak > 0 else None
        roofline["alone_avg_launch_ms"] = a_ms
        roofline["alone_pivots_per_s"] = alone["pivots"] / (alone["dev_ms"] * 1e-3)

    # ---- parity, in the run: the state after 200 pivots against the CPU side ----
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    parity = {}
    Kp = 200
    lp.set_block(args.block)
    lp.fill_synthetic(SEED)
    lp.solve(Kp)
    ct, cg = lp.checksum()
    ct, cg = all_sum_u64(ct), all_sum_u64(cg)
    e2b_K = lp.download(want_tab=False)["eq2bv"]
    gold_path = os.path.join(ROOT, "tests", "golden", "c3_checkpoints.json")
    if os.path.exists(gold_path):
        gold = json.load(open(gold_path))
        if (gold["m"], gold["n"], gold["seed"]) == (m, n, SEED):
            ref = gold["reference"][str(Kp)]
            wts = np.arange(1, m + 1, dtype=np.int64)
            parity["checksum_matches_reference"] = bool(
                ct == ref["tab"] and cg == ref["tgtf"] and int(e2b_K.astype(np.int64).dot(wts)) == ref["eq2bv_sum"])
            parity["reference_checkpoint"] = (f"tableau + objective-row checksums and basis after {Kp} pivots, computed "
                                              "by the unmodified reference (tests/golden/c3_checkpoints.json)")
    if world > 1:  # the shards against ONE GPU running the same LP
        ok1 = True
        if rank == 0:
            one = ctx.large_lp(m, Ccols)
            one.set_block(args.block)
            one.fill_synthetic(SEED)
            one.solve(Kp)
            o1 = one.checksum()
            ok1 = (o1[0] == ct and o1[1] == cg and np.array_equal(one.download(want_tab=False)["eq2bv"], e2b_K))
            one.close()
        parity["checksum_matches_1gpu"] = bcast_flag(ok1)
    if not args.no_cpu:
        okb = True
        Kc = args.cpu_pivots
        lp.fill_synthetic(SEED)
        lp.solve(Kc)
        g = lp.download(want_tab=False, log_cap=Kc)
        if rank == 0:
            from xpoly_b200.synth import dense_lp
            leq_h, tg_h = dense_lp(SEED, m, n)
            rate, kind, info = cpu_pivot_rate(leq_h, tg_h, Kc, prefer_ref=False)
            okb = bool(np.array_equal(g["eq2bv"], info["eq2bv"]))
            cpu_line = {"value": rate, "unit": "pivots/s", "cores": 1, "kind": kind,
                        "sample": f"{Kc} pivots of the same {m}x{Ccols} LP (single thread; the reference has no "
                                  "intra-LP parallelism); solve loop timed inside the checker"}
        parity["basis_matches_oracle"] = bcast_flag(okb)
    parity_ok = all(v for k, v in parity.items() if isinstance(v, bool))

    line = {
        "metric": "pivots/s", "value": value, "unit": "pivots/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["dev_ms"] / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"c3: dense FP64 LP, tableau {m}x{Ccols}, simplex iterations "
                               "under the reference pivot rule from the slack basis",
                   "placement": "HBM-resident" + (f", column-sharded over {world} GPUs" if world > 1 else ""),
                   "l2": "inputs (1 GiB) larger than L2 (126 MB): no flush between iterations",
                   "pivots_per_step": P, "pivots_per_tableau_pass": k_eff, "pricing_window": lp.window},
        "gpu_launches": int(main["launches"]), "clocks": clocks, "roofline": roofline, "parity": parity,
    }
    details = {"details_of": "bench.py secondary legs (the main line follows)", "n_gpus": world}

    if world == 1:
        # ---- e2e: the C-ABI call the reference arm's TwoStageMethod(leq, tgtf) corresponds to --
        # xp_six_two_stage_f64_large with PINNED host buffers: leq goes up (half the slack form),
        # slack form + P pivots on the device, O(C) comes down
        from xpoly_b200.synth import dense_lp
        leq_h, tg_h = dense_lp(SEED, m, n)
        hp = C.c_void_p()
        ctx.check(lib.xp_host_alloc(ctx._h, C.c_size_t(leq_h.nbytes), C.byref(hp)))
        h_leq = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=leq_h.shape)
        h_leq[...] = leq_h
        ctx.set_block(args.block)
        if args.window is not None:
            ctx.set_window(args.window)
        stv = C.c_int32(0)
        mv, ssol, otg = np.zeros(1), np.zeros(Ccols), np.zeros(Ccols)
        e2b, its, pvs = np.zeros(m, dtype=np.int32), np.zeros(1, dtype=np.uint32), np.zeros(1, dtype=np.uint32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)

        def e2e_step():
            t0 = time.perf_counter()
            ctx.check(lib.xp_six_two_stage_f64_large(ctx._h, m, n, hp, p(tg_h), C.c_uint32(P), 0, C.byref(stv),
                                                     p(mv), p(ssol), p(otg), p(e2b), p(its), p(pvs)))
            return time.perf_counter() - t0, int(its[0])
        e2e_step()
        ts, itn = [], 0
        for _ in range(max(3, min(args.steps, 8))):
            dt, it = e2e_step()
            ts.append(dt)
            itn += it
        line["e2e"] = {"value": itn / sum(ts), "unit": "pivots/s",
                       "h2d_bytes_per_step": int(leq_h.nbytes + tg_h.nbytes),
                       "d2h_bytes_per_step": int(2 * Ccols * 8 + m * 4 + 64),
                       "ms_per_step": 1000.0 * sum(ts) / len(ts), "pivots_per_step": P,
                       "api": "xp_six_two_stage_f64_large (pinned host leq; slack form on the device)",
                       "basis_matches_device_run": bool(np.array_equal(e2b, e2b_K)) if P == Kp else None}
        ctx.check(lib.xp_host_free(ctx._h, hp))
        del h_leq
        # the kernel-level entry (replaces the private solveSlackForm: whole tableau up and down)
        tab_bytes = m * Ccols * 8
        ctx.check(lib.xp_host_alloc(ctx._h, C.c_size_t(tab_bytes), C.byref(hp)))
        lp.set_block(args.block)
        lp.fill_synthetic(SEED)
        st0 = lp.download(want_tab=False)
        ctx.check(lib.xp_lp_f64_download(lp._h, hp, None, None, None, None, None, None, None, None, None, 0))
        lp.close()
        tg, nv, bvs = st0["tgtf"].copy(), st0["nvset"].copy(), st0["bvset"].copy()
        b2e, e2b2 = st0["bv2eq"].copy(), st0["eq2bv"].copy()
        sol = np.zeros(Ccols)

        def slack_step():
            t0 = time.perf_counter()
            ctx.check(lib.xp_six_slack_f64(ctx._h, hp, p(tg), m, Ccols, p(nv), p(bvs), p(b2e), p(e2b2), None, None,
                                           C.c_uint32(P), 0, p(mv), p(sol), p(its), None, 0))
            return time.perf_counter() - t0, int(its[0])
        slack_step()
        ts, itn = [], 0
        for _ in range(2):
            dt, it = slack_step()
            ts.append(dt)
            itn += it
        line["e2e_slack"] = {"value": itn / sum(ts), "unit": "pivots/s", "bytes_each_way": int(tab_bytes),
                             "api": "xp_six_slack_f64 (whole tableau up and down per step)"}
        ctx.check(lib.xp_host_free(ctx._h, hp))
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_line
        if not args.no_batched:
            details["batched"] = run_batched(ctx, xp, torch, dev, with_cpu=not args.no_cpu)
            details.update(run_exact_and_bnb(ctx, xp, torch, dev, with_cpu=not args.no_cpu))
    else:
        # ---- e2e at N GPUs: the LP starts in (pinned) host memory as the caller's leq / tgtf; every
        # rank uploads the columns of leq that fall into its slice (xp_lp_f64_upload_leq on the sharded
        # handle: 2-D copy over its own PCIe link, slack columns generated on the device), the ranks
        # solve together, O(C) comes down; all copies inside the timed region, wall clock between
        # barriers, max over ranks
        from xpoly_b200.synth import dense_lp
        leq_h, tg_h = dense_lp(SEED, m, n)
        hp_in = C.c_void_p()
        ctx.check(lib.xp_host_alloc(ctx._h, C.c_size_t(leq_h.nbytes), C.byref(hp_in)))
        h_leq = np.ctypeslib.as_array(C.cast(hp_in, C.POINTER(C.c_double)), shape=leq_h.shape)
        h_leq[...] = leq_h
        lp.set_block(args.block)
        tg1 = np.zeros(Ccols)
        e2b1 = np.zeros(m, dtype=np.int32)
        maxv, sol = np.zeros(1), np.zeros(Ccols)
        iters = np.zeros(1, dtype=np.uint32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)

        def e2e_step():
            barrier()
            t0 = time.perf_counter()
            ctx.check(lib.xp_lp_f64_upload_leq(lp._h, hp_in, p(tg_h), n))
            st = ctx.check(lib.xp_lp_f64_solve(lp._h, C.c_uint32(P), 0))
            ctx.check(lib.xp_lp_f64_download(lp._h, None, p(tg1), None, None, None, p(e2b1), p(maxv), p(sol), p(iters),
                                             None, 0))
            barrier()
            return time.perf_counter() - t0, int(iters[0]), st
        e2e_step()
        ts, its = [], 0
        for _ in range(max(3, min(args.steps, 8))):
            dt, it, _st = e2e_step()
            ts.append(dt)
            its += it
        tt = torch.tensor([sum(ts)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        a_lo, a_hi = min(lp.col0, n), min(lp.col0 + local_cols, n)
        up = torch.tensor([float(m * max(a_hi - a_lo, 0) * 8 + m * 8 + (n + 1) * 8)], dtype=torch.float64, device=dev)
        dist.all_reduce(up)
        line["e2e"] = {"value": its / float(tt.item()), "unit": "pivots/s",
                       "h2d_bytes_per_step": int(up.item()),
                       "d2h_bytes_per_step": int(world * (2 * Ccols * 8 + m * 4 + 64)),
                       "ms_per_step": 1000.0 * float(tt.item()) / len(ts), "pivots_per_step": P,
                       "api": "xp_lp_f64_upload_leq + xp_lp_f64_solve + xp_lp_f64_download on the column-sharded "
                              "handle (pinned host leq; every rank uploads the columns of its slice over its own "
                              "PCIe link, slack form on the device)",
                       "basis_matches_device_run": bool(np.array_equal(e2b1, e2b_K)) if P == Kp else None}
        ctx.check(lib.xp_host_free(ctx._h, hp_in))
        del h_leq
        barrier()
        lp.close()
        barrier()
        if not args.no_batched:
            details["batched"] = run_batched(ctx, xp, torch, dev, with_cpu=False, rank=rank,
                                             world=world, dist=dist)
            details.update(run_exact_and_bnb(ctx, xp, torch, dev, rank=rank, world=world, dist=dist,
                                             with_cpu=False))
    # flat copies of the secondary figures in the main line (the driver keeps one line)
    for k, key in (("batched", "c2_LPs_per_s"), ("exact", "c4_exact_LPs_per_s"), ("bnb", "c5_node_LPs_per_s"),
                   ("has_solution", "has_solution_queries_per_s")):
        if k in details:
            line[key] = details[k]["value"]
    if rank == 0:
        if len(details) > 2:
            print(json.dumps(details), flush=True)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0 if parity_ok else 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pivots", type=int, default=200, help="simplex iterations per step (SURVEY 8d: K = 200)")
    ap.add_argument("--block", type=int, default=0, help="pivots per tableau pass (0 = automatic)")
    ap.add_argument("--rank1-pivots", type=int, default=40)
    ap.add_argument("--m", type=int, default=M_ROWS)
    ap.add_argument("--n", type=int, default=N_VARS)
    ap.add_argument("--cpu-pivots", type=int, default=12)
    ap.add_argument("--ref-pivots", type=int, default=3)
    ap.add_argument("--port", action="store_true", help="reference arm: force the oracle port")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--window", type=int, default=None, help="pricing window of the panel kernel (default: automatic; -1 off)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
