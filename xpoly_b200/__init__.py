"""xpoly_b200: ctypes binding of the B200-native simplex hot path (C ABI in
include/xpoly_b200.h).  Python is only the test / bench harness language here;
the product is libxpoly_b200.so plus the C++ adaptor in xpoly_b200/host/.

There is no CPU fallback: if the shared library is missing, or no CUDA device
is usable, the calls below raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxpoly_b200.so")

SIX_SUCC, SIX_UNBOUND, SIX_NO_PRI_FEASIBLE_SOL, SIX_OPTIMAL_IS_INFEASIBLE, SIX_TIME_OUT = range(5)
IP_SUCC, IP_UNBOUND, IP_NO_PRI_FEASIBLE_SOL, IP_NO_BETTER_THAN_BEST_SOL = range(4)
ERR_CUDA, ERR_BAD_ARG, ERR_TOO_LARGE, ERR_OVERFLOW, ERR_PEER = -1, -2, -3, -4, -5
MAX_RANKS, PEER_HANDLE_BYTES = 8, 64
ERR_REFERENCE_UB = -100
RULE_REFERENCE = 0
NO_ITER_LIMIT = 0xFFFFFFFF

_vp = C.c_void_p
_lib = None


class XpolyError(RuntimeError):
    pass


def lib():
    """The loaded C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XpolyError(
                f"{LIB_PATH} is missing: run `python -m xpoly_b200.build` "
                "(xpoly_b200 has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.xp_last_error.restype = C.c_char_p
        _lib.xp_version.restype = C.c_char_p
        _lib.xp_ctx_launch_count.restype = C.c_uint64
        _lib.xp_ctx_last_kernel_ms.restype = C.c_float
        _lib.xp_ctx_stream.restype = _vp
        for name in ("xp_ctx_destroy", "xp_lp_f64_destroy"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One xp_ctx: a CUDA device + stream + reusable device scratch."""

    def __init__(self, device=0):
        self._h = _vp()
        rc = lib().xp_ctx_create(int(device), C.byref(self._h))
        if rc != 0:
            msg = lib().xp_last_error(self._h).decode() if self._h else "xp_ctx_create failed"
            if self._h:
                lib().xp_ctx_destroy(self._h)
                self._h = _vp()
            raise XpolyError(msg)

    def close(self):
        if self._h:
            lib().xp_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        """Negative codes are errors; statuses (>= 0) pass through."""
        if rc < 0 and rc not in (ERR_REFERENCE_UB, ERR_OVERFLOW, ERR_TOO_LARGE):
            raise XpolyError(f"xpoly_b200 error {rc}: {lib().xp_last_error(self._h).decode()}")
        return rc

    @property
    def launches(self):
        return int(lib().xp_ctx_launch_count(self._h))

    @property
    def last_kernel_ms(self):
        return float(lib().xp_ctx_last_kernel_ms(self._h))

    @property
    def stream(self):
        return lib().xp_ctx_stream(self._h)

    # ---- kernel level: SIX::solveSlackForm on a caller-built slack form ----
    def six_slack_f64(self, tab, tgtf, nvset, bvset, bv2eq, eq2bv, max_iter=NO_ITER_LIMIT,
                      vc_diag=None, vc_rhs=None, log_cap=0):
        """In place on copies of the inputs; returns a dict of the outputs."""
        tab = _f64(tab).copy()
        tgtf = _f64(tgtf).copy()
        m, Cc = tab.shape
        nvset = np.ascontiguousarray(nvset, dtype=np.uint8).copy()
        bvset = np.ascontiguousarray(bvset, dtype=np.uint8).copy()
        bv2eq = np.ascontiguousarray(bv2eq, dtype=np.int32).copy()
        eq2bv = np.ascontiguousarray(eq2bv, dtype=np.int32).copy()
        maxv = np.zeros(1)
        sol = np.zeros(Cc)
        iters = np.zeros(1, dtype=np.uint32)
        log = np.zeros((max(log_cap, 1), 3), dtype=np.int32)
        st = self.check(lib().xp_six_slack_f64(
            self._h, _p(tab), _p(tgtf), m, Cc, _p(nvset), _p(bvset), _p(bv2eq), _p(eq2bv),
            _p(None if vc_diag is None else _f64(vc_diag)),
            _p(None if vc_rhs is None else _f64(vc_rhs)), C.c_uint32(max_iter), RULE_REFERENCE,
            _p(maxv), _p(sol), _p(iters), _p(log) if log_cap else None, C.c_uint32(log_cap)))
        n_it = int(iters[0])
        return dict(status=st, tab=tab, tgtf=tgtf, nvset=nvset, bvset=bvset, bv2eq=bv2eq,
                    eq2bv=eq2bv, maxv=maxv, sol=sol, iters=n_it, log=log[:min(n_it, log_cap)])

    def set_block(self, k):
        """Pivots per pass over the tableau for six_slack_f64 (0 = automatic)."""
        self.check(lib().xp_ctx_set_block(self._h, int(k)))

    def set_window(self, w):
        """Pricing window for six_slack_f64 / two_stage_f64_large: 0 automatic, < 0 off, > 0 forced."""
        self.check(lib().xp_ctx_set_window(self._h, int(w)))

    def last_lp_checksum(self):
        """(tableau, objective row) checksums of what the last large two-stage / slack call left on the device."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.check(lib().xp_ctx_last_lp_checksum(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def large_lp(self, m, Cc, rank=0, nranks=1):
        return LargeLP(self, m, Cc, rank, nranks)


class LargeLP:
    """Device-resident FP64 slack-form LP (xp_lp_f64 handle); with nranks > 1 this
    is rank `rank`'s column shard (see xpoly_b200/sharded.py for the handshake)."""

    def __init__(self, ctx, m, Cc, rank=0, nranks=1):
        self.ctx, self.m, self.C, self.rank, self.nranks = ctx, m, Cc, rank, nranks
        self._h = _vp()
        if nranks == 1:
            ctx.check(lib().xp_lp_f64_create(ctx._h, m, Cc, C.byref(self._h)))
        else:
            ctx.check(lib().xp_lp_f64_create_sharded(ctx._h, m, Cc, rank, nranks, C.byref(self._h)))
        c0, nc = C.c_int(0), C.c_int(0)
        ctx.check(lib().xp_lp_f64_local_cols(self._h, C.byref(c0), C.byref(nc)))
        self.col0, self.local_cols = c0.value, nc.value

    def set_block(self, k):
        self.ctx.check(lib().xp_lp_f64_set_block(self._h, int(k)))

    def set_window(self, w):
        """Pricing window of the panel kernel: 0 automatic, < 0 off, > 0 forced width."""
        self.ctx.check(lib().xp_lp_f64_set_window(self._h, int(w)))

    @property
    def window(self):
        return int(lib().xp_lp_f64_window(self._h))

    def peer_handle(self):
        buf = np.zeros(PEER_HANDLE_BYTES, dtype=np.uint8)
        self.ctx.check(lib().xp_lp_f64_peer_handle(self._h, _p(buf)))
        return buf

    def peer_attach(self, handles):
        """handles: uint8 [nranks, PEER_HANDLE_BYTES] in rank order (all-gathered)."""
        h = np.ascontiguousarray(handles, dtype=np.uint8).reshape(self.nranks, PEER_HANDLE_BYTES)
        self.ctx.check(lib().xp_lp_f64_peer_attach(self._h, _p(h)))

    def peer_attach_local(self, shards):
        arr = (_vp * self.nranks)(*[s._h for s in shards])
        self.ctx.check(lib().xp_lp_f64_peer_attach_local(self._h, arr))

    def close(self):
        if self._h:
            lib().xp_lp_f64_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, tab, tgtf, nvset, bvset, bv2eq, eq2bv, vc_diag=None, vc_rhs=None):
        self.ctx.check(lib().xp_lp_f64_upload(
            self._h, _p(_f64(tab)), _p(_f64(tgtf)),
            _p(np.ascontiguousarray(nvset, dtype=np.uint8)),
            _p(np.ascontiguousarray(bvset, dtype=np.uint8)),
            _p(np.ascontiguousarray(bv2eq, dtype=np.int32)),
            _p(np.ascontiguousarray(eq2bv, dtype=np.int32)),
            _p(None if vc_diag is None else _f64(vc_diag)),
            _p(None if vc_rhs is None else _f64(vc_rhs))))

    def upload_leq(self, leq, tgtf):
        leq = _f64(leq)
        n = leq.shape[1] - 1
        self.ctx.check(lib().xp_lp_f64_upload_leq(self._h, _p(leq), _p(_f64(tgtf)), n))

    def fill_synthetic(self, seed):
        self.ctx.check(lib().xp_lp_f64_fill_synthetic(self._h, C.c_uint64(seed)))

    def solve(self, max_iter=NO_ITER_LIMIT):
        return self.ctx.check(lib().xp_lp_f64_solve(self._h, C.c_uint32(max_iter), RULE_REFERENCE))

    def checksum(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.ctx.check(lib().xp_lp_f64_checksum(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def download(self, want_tab=True, log_cap=0):
        m, Cc = self.m, self.C
        n = Cc - 1
        tab = np.zeros((m, Cc)) if want_tab else None
        tgtf = np.zeros(Cc)
        nvset = np.zeros(n, dtype=np.uint8)
        bvset = np.zeros(n, dtype=np.uint8)
        bv2eq = np.zeros(n, dtype=np.int32)
        eq2bv = np.zeros(m, dtype=np.int32)
        maxv = np.zeros(1)
        sol = np.zeros(Cc)
        iters = np.zeros(1, dtype=np.uint32)
        log = np.zeros((max(log_cap, 1), 3), dtype=np.int32)
        self.ctx.check(lib().xp_lp_f64_download(
            self._h, _p(tab), _p(tgtf), _p(nvset), _p(bvset), _p(bv2eq), _p(eq2bv), _p(maxv),
            _p(sol), _p(iters), _p(log) if log_cap else None, C.c_uint32(log_cap)))
        n_it = int(iters[0])
        return dict(tab=tab, tgtf=tgtf, nvset=nvset, bvset=bvset, bv2eq=bv2eq, eq2bv=eq2bv,
                    maxv=maxv, sol=sol, iters=n_it, log=log[:min(n_it, log_cap)])


def slack_form(leq, tgtf):
    """Host-side SIX::slack + identity basis (lpsol.h:1405-1433, :1821-1841) for a
    normalised LP: returns tab [A | I | b], tgtf, nvset, bvset, bv2eq, eq2bv."""
    leq = _f64(leq)
    m, n1 = leq.shape
    n = n1 - 1
    Cc = n + m + 1
    tab = np.zeros((m, Cc))
    tab[:, :n] = leq[:, :n]
    tab[:, n:n + m] = np.eye(m)
    tab[:, Cc - 1] = leq[:, n]
    tg = np.zeros(Cc)
    tg[:n] = np.asarray(tgtf, dtype=np.float64)[:n]
    tg[Cc - 1] = np.asarray(tgtf, dtype=np.float64)[n]
    nvset = np.zeros(Cc - 1, dtype=np.uint8)
    nvset[:n] = 1
    bvset = (1 - nvset).astype(np.uint8)
    bv2eq = np.full(Cc - 1, -1, dtype=np.int32)
    bv2eq[n:] = np.arange(m, dtype=np.int32)
    eq2bv = (n + np.arange(m)).astype(np.int32)
    return tab, tg, nvset, bvset, bv2eq, eq2bv


def _two_stage_f64_batch(self, leq, tgtf, max_iter=NO_ITER_LIMIT, want=("status", "maxv", "slack_sol",
                                                                       "tgtf", "eq2bv", "iters",
                                                                       "pivots")):
    """SIX::TwoStageMethod for a uniform batch: leq [B, m, n+1], tgtf [B, n+1]."""
    leq = _f64(leq)
    tgtf = _f64(tgtf)
    B, m, n1 = leq.shape
    n = n1 - 1
    ldo = n + m + 1
    out = dict(
        status=np.zeros(B, dtype=np.int32) if "status" in want else None,
        maxv=np.zeros(B) if "maxv" in want else None,
        slack_sol=np.zeros((B, ldo)) if "slack_sol" in want else None,
        tgtf=np.zeros((B, ldo)) if "tgtf" in want else None,
        eq2bv=np.zeros((B, m), dtype=np.int32) if "eq2bv" in want else None,
        iters=np.zeros(B, dtype=np.uint32) if "iters" in want else None,
        pivots=np.zeros(B, dtype=np.uint32) if "pivots" in want else None)
    self.check(lib().xp_six_two_stage_f64_batch(
        self._h, B, m, n, _p(leq), _p(tgtf), C.c_uint32(max_iter), RULE_REFERENCE,
        _p(out["status"]), _p(out["maxv"]), _p(out["slack_sol"]), _p(out["tgtf"]),
        _p(out["eq2bv"]), _p(out["iters"]), _p(out["pivots"])))
    return out


def _two_stage_f64_ragged(self, lps, max_iter=NO_ITER_LIMIT):
    """lps: list of (leq [m, n+1], tgtf [n+1]) with arbitrary per-LP shapes."""
    B = len(lps)
    ms = np.array([l.shape[0] for l, _ in lps], dtype=np.int32)
    ns = np.array([l.shape[1] - 1 for l, _ in lps], dtype=np.int32)
    leq_pool = np.concatenate([_f64(l).ravel() for l, _ in lps])
    tg_pool = np.concatenate([_f64(t).ravel() for _, t in lps])
    leq_off = np.zeros(B, dtype=np.int64)
    tg_off = np.zeros(B, dtype=np.int64)
    leq_off[1:] = np.cumsum(ms[:-1].astype(np.int64) * (ns[:-1] + 1))
    tg_off[1:] = np.cumsum(ns[:-1].astype(np.int64) + 1)
    ldo = int((ms + ns).max()) + 1
    ldm = int(ms.max())
    out = dict(status=np.zeros(B, dtype=np.int32), maxv=np.zeros(B), slack_sol=np.zeros((B, ldo)),
               tgtf=np.zeros((B, ldo)), eq2bv=np.zeros((B, ldm), dtype=np.int32),
               iters=np.zeros(B, dtype=np.uint32), pivots=np.zeros(B, dtype=np.uint32))
    self.check(lib().xp_six_two_stage_f64_ragged(
        self._h, B, _p(ms), _p(ns), _p(leq_off), _p(tg_off), _p(leq_pool),
        C.c_size_t(leq_pool.size), _p(tg_pool), C.c_size_t(tg_pool.size), C.c_uint32(max_iter),
        RULE_REFERENCE, ldo, ldm, _p(out["status"]), _p(out["maxv"]), _p(out["slack_sol"]),
        _p(out["tgtf"]), _p(out["eq2bv"]), _p(out["iters"]), _p(out["pivots"])))
    out["ms"], out["ns"] = ms, ns
    return out


def _two_stage_f64_large(self, leq, tgtf, max_iter=NO_ITER_LIMIT):
    """SIX::TwoStageMethod for one LP on the HBM-resident path (phase 1 on the device)."""
    leq, tgtf = _f64(leq), _f64(tgtf)
    m, n = leq.shape[0], leq.shape[1] - 1
    Cc = n + m + 1
    st = C.c_int32(0)
    out = dict(maxv=np.zeros(1), slack_sol=np.zeros(Cc), tgtf=np.zeros(Cc),
               eq2bv=np.zeros(m, dtype=np.int32), iters=np.zeros(1, dtype=np.uint32),
               pivots=np.zeros(1, dtype=np.uint32))
    self.check(lib().xp_six_two_stage_f64_large(
        self._h, m, n, _p(leq), _p(tgtf), C.c_uint32(max_iter), RULE_REFERENCE, C.byref(st),
        _p(out["maxv"]), _p(out["slack_sol"]), _p(out["tgtf"]), _p(out["eq2bv"]), _p(out["iters"]),
        _p(out["pivots"])))
    out["status"] = st.value
    out["iters"], out["pivots"] = int(out["iters"][0]), int(out["pivots"][0])
    return out


Context.two_stage_f64_large = _two_stage_f64_large
Context.two_stage_f64_batch = _two_stage_f64_batch
Context.two_stage_f64_ragged = _two_stage_f64_ragged


def _two_stage_i64_batch(self, leq, tgtf, max_iter=NO_ITER_LIMIT):
    """Exact (fraction-free) TwoStageMethod for a uniform batch of integer LPs:
    leq [B, m, n+1], tgtf [B, n+1] (integer valued).  Outputs are reduced
    num/den int64 pairs."""
    leq = np.ascontiguousarray(leq, dtype=np.int64)
    tgtf = np.ascontiguousarray(tgtf, dtype=np.int64)
    B, m, n1 = leq.shape
    n = n1 - 1
    ldo = n + m + 1
    out = dict(status=np.zeros(B, dtype=np.int32), maxv=np.zeros((B, 2), dtype=np.int64),
               sol_num=np.zeros((B, ldo), dtype=np.int64), sol_den=np.zeros((B, ldo), dtype=np.int64),
               tgtf_num=np.zeros((B, ldo), dtype=np.int64), tgtf_den=np.zeros((B, ldo), dtype=np.int64),
               eq2bv=np.zeros((B, m), dtype=np.int32), iters=np.zeros(B, dtype=np.uint32),
               pivots=np.zeros(B, dtype=np.uint32))
    self.check(lib().xp_six_two_stage_i64_batch(
        self._h, B, m, n, _p(leq), _p(tgtf), C.c_uint32(max_iter), RULE_REFERENCE,
        _p(out["status"]), _p(out["maxv"]), _p(out["sol_num"]), _p(out["sol_den"]),
        _p(out["tgtf_num"]), _p(out["tgtf_den"]), _p(out["eq2bv"]), _p(out["iters"]),
        _p(out["pivots"])))
    return out


def _two_stage_i64_ragged(self, lps, max_iter=NO_ITER_LIMIT):
    B = len(lps)
    ms = np.array([l.shape[0] for l, _ in lps], dtype=np.int32)
    ns = np.array([l.shape[1] - 1 for l, _ in lps], dtype=np.int32)
    leq_pool = np.concatenate([np.asarray(l, dtype=np.int64).ravel() for l, _ in lps])
    tg_pool = np.concatenate([np.asarray(t, dtype=np.int64).ravel() for _, t in lps])
    leq_off = np.zeros(B, dtype=np.int64)
    tg_off = np.zeros(B, dtype=np.int64)
    leq_off[1:] = np.cumsum(ms[:-1].astype(np.int64) * (ns[:-1] + 1))
    tg_off[1:] = np.cumsum(ns[:-1].astype(np.int64) + 1)
    ldo = int((ms + ns).max()) + 1
    ldm = int(ms.max())
    out = dict(status=np.zeros(B, dtype=np.int32), maxv=np.zeros((B, 2), dtype=np.int64),
               sol_num=np.zeros((B, ldo), dtype=np.int64), sol_den=np.zeros((B, ldo), dtype=np.int64),
               tgtf_num=np.zeros((B, ldo), dtype=np.int64), tgtf_den=np.zeros((B, ldo), dtype=np.int64),
               eq2bv=np.zeros((B, ldm), dtype=np.int32), iters=np.zeros(B, dtype=np.uint32),
               pivots=np.zeros(B, dtype=np.uint32))
    self.check(lib().xp_six_two_stage_i64_ragged(
        self._h, B, _p(ms), _p(ns), _p(leq_off), _p(tg_off), _p(leq_pool),
        C.c_size_t(leq_pool.size), _p(tg_pool), C.c_size_t(tg_pool.size), C.c_uint32(max_iter),
        RULE_REFERENCE, ldo, ldm, _p(out["status"]), _p(out["maxv"]), _p(out["sol_num"]),
        _p(out["sol_den"]), _p(out["tgtf_num"]), _p(out["tgtf_den"]), _p(out["eq2bv"]),
        _p(out["iters"]), _p(out["pivots"])))
    out["ms"], out["ns"] = ms, ns
    return out


Context.two_stage_i64_batch = _two_stage_i64_batch
Context.two_stage_i64_ragged = _two_stage_i64_ragged


# --------------------------------------------------------------- entry level
def _rat(a):
    """int (..., 2) array of num/den pairs, or an integer array (den = 1)."""
    a = np.asarray(a)
    if a.dtype != np.int32 or a.shape[-1:] != (2,):
        out = np.empty(a.shape + (2,), dtype=np.int32)
        out[..., 0] = a
        out[..., 1] = 1
        a = out
    return np.ascontiguousarray(a, dtype=np.int32)


def _six_solve(self, kind, is_min, leq, tgtf, vc=None, eq=None, max_iter=NO_ITER_LIMIT):
    """SIX::maxm / minm.  kind 'f64' (float64 arrays) or 'rat' (int32 num/den pairs)."""
    if kind == "f64":
        conv, v = _f64, np.zeros(1)
    else:
        conv, v = _rat, np.zeros(2, dtype=np.int32)
    leq = None if leq is None else conv(leq)
    eq = None if eq is None else conv(eq)
    tgtf = conv(tgtf)
    vc = None if vc is None else conv(vc)
    n = tgtf.shape[0] - 1
    m = 0 if leq is None else leq.shape[0]
    k = 0 if eq is None else eq.shape[0]
    sol = np.zeros((n + 1,) + tgtf.shape[1:], dtype=tgtf.dtype)
    # exactly the documented capacity (m + 2k for maxm, 2n for minm) plus a canary word the
    # library must never touch
    cap = 2 * n if is_min else m + 2 * k
    e2b = np.full(cap + 1, -1, dtype=np.int32)
    e2b[cap] = 0x5AFE5AFE
    fn = getattr(lib(), f"xp_six_{'minm' if is_min else 'maxm'}_{kind}")
    st = self.check(fn(self._h, m, n, _p(tgtf), _p(vc), k, _p(eq), _p(leq), C.c_uint32(max_iter),
                       _p(v), _p(sol), _p(e2b)))
    if e2b[cap] != 0x5AFE5AFE:
        raise RuntimeError("xp_six_*: eq2bv_out written past its documented capacity")
    return dict(status=st, v=v, sol=sol, eq2bv=e2b[:cap])


def _six_solve_batch(self, kind, is_min, leq, tgtf, max_iter=NO_ITER_LIMIT):
    conv = _f64 if kind == "f64" else _rat
    leq, tgtf = conv(leq), conv(tgtf)
    B, m, n1 = leq.shape[:3]
    status = np.zeros(B, dtype=np.int32)
    v = np.zeros((B,) + tgtf.shape[2:], dtype=tgtf.dtype)
    sol = np.zeros((B, n1) + tgtf.shape[2:], dtype=tgtf.dtype)
    fn = getattr(lib(), f"xp_six_solve_{kind}_batch")
    self.check(fn(self._h, int(is_min), B, m, n1 - 1, _p(tgtf), _p(leq), C.c_uint32(max_iter),
                  _p(status), _p(v), _p(sol)))
    return dict(status=status, v=v, sol=sol)


def _mip_solve(self, kind, is_min, is_bin, leq, tgtf, eq=None):
    if kind == "f64":
        conv, v = _f64, np.zeros(1)
    else:
        conv, v = _rat, np.zeros(2, dtype=np.int32)
    leq = None if leq is None else conv(leq)
    eq = None if eq is None else conv(eq)
    tgtf = conv(tgtf)
    n = tgtf.shape[0] - 1
    m = 0 if leq is None else leq.shape[0]
    k = 0 if eq is None else eq.shape[0]
    sol = np.zeros((n + 1,) + tgtf.shape[1:], dtype=tgtf.dtype)
    nodes = C.c_int32(0)
    fn = getattr(lib(), f"xp_mip_solve_{kind}")
    st = self.check(fn(self._h, int(is_min), int(is_bin), m, n, _p(tgtf), k, _p(eq), _p(leq),
                       _p(v), _p(sol), C.byref(nodes)))
    return dict(status=st, v=v, sol=sol, nodes=nodes.value)


def _mip_solve_rat_ri(self, is_min, is_bin, leq, tgtf, indicator, eq=None):
    """MIP<RMat,Rational> with rational_indicator: n+1 flags, non-zero = may stay rational."""
    leq = None if leq is None else _rat(leq)
    eq = None if eq is None else _rat(eq)
    tgtf = _rat(tgtf)
    n = tgtf.shape[0] - 1
    m = 0 if leq is None else leq.shape[0]
    k = 0 if eq is None else eq.shape[0]
    v = np.zeros(2, dtype=np.int32)
    sol = np.zeros((n + 1, 2), dtype=np.int32)
    nodes = C.c_int32(0)
    ind = np.ascontiguousarray(indicator, dtype=np.uint8)
    st = self.check(lib().xp_mip_solve_rat_ri(self._h, int(is_min), int(is_bin), m, n, _p(tgtf), k, _p(eq),
                                              _p(leq), _p(ind), _p(v), _p(sol), C.byref(nodes)))
    return dict(status=st, v=v, sol=sol, nodes=nodes.value)


def _mip_solve_rat_batch(self, is_min, is_bin, leq, tgtf):
    leq, tgtf = _rat(leq), _rat(tgtf)
    B, m, n1 = leq.shape[:3]
    status = np.zeros(B, dtype=np.int32)
    v = np.zeros((B, 2), dtype=np.int32)
    sol = np.zeros((B, n1, 2), dtype=np.int32)
    nodes = np.zeros(B, dtype=np.int32)
    self.check(lib().xp_mip_solve_rat_batch(self._h, int(is_min), int(is_bin), B, m, n1 - 1,
                                            _p(tgtf), _p(leq), _p(status), _p(v), _p(sol),
                                            _p(nodes)))
    return dict(status=status, v=v, sol=sol, nodes=nodes)


def _has_solution_batch(self, leq, is_int=True, is_unique=True):
    leq = _rat(leq)
    B, m, n1 = leq.shape[:3]
    res = np.zeros(B, dtype=np.int32)
    self.check(lib().xp_has_solution_rat_batch(self._h, B, m, n1 - 1, _p(leq), int(is_int),
                                               int(is_unique), _p(res)))
    return res


def _has_solution_ragged(self, systems, is_int=True, is_unique=True):
    """systems: list of (leq or None, eq or None), integer or (…, 2) int32 rational arrays with
    n+1 columns each.  Returns int32 results (1 / 0 / negative per-system code)."""
    B = len(systems)
    ns = np.zeros(B, dtype=np.int32)
    ms = np.zeros(B, dtype=np.int32)
    ks = np.zeros(B, dtype=np.int32)
    lo = np.zeros(B, dtype=np.int64)
    eo = np.zeros(B, dtype=np.int64)
    lp, ep = [], []
    ll = el = 0
    for b, (leq, eq) in enumerate(systems):
        leq = None if leq is None or len(leq) == 0 else _rat(leq)
        eq = None if eq is None or len(eq) == 0 else _rat(eq)
        ns[b] = (leq if leq is not None else eq).shape[1] - 1
        if leq is not None:
            ms[b], lo[b] = leq.shape[0], ll
            lp.append(leq.reshape(-1, 2))
            ll += leq.shape[0] * leq.shape[1]
        if eq is not None:
            ks[b], eo[b] = eq.shape[0], el
            ep.append(eq.reshape(-1, 2))
            el += eq.shape[0] * eq.shape[1]
    lpool = np.ascontiguousarray(np.concatenate(lp)) if lp else np.zeros((1, 2), dtype=np.int32)
    epool = np.ascontiguousarray(np.concatenate(ep)) if ep else np.zeros((1, 2), dtype=np.int32)
    res = np.zeros(B, dtype=np.int32)
    self.check(lib().xp_has_solution_rat_ragged(self._h, B, _p(ns), _p(ms), _p(lo), _p(lpool),
                                                C.c_size_t(ll), _p(ks), _p(eo), _p(epool),
                                                C.c_size_t(el), int(is_int), int(is_unique),
                                                _p(res)))
    return res


Context.has_solution_ragged = _has_solution_ragged
Context.six_solve = _six_solve
Context.six_solve_batch = _six_solve_batch
Context.mip_solve = _mip_solve
Context.mip_solve_rat_ri = _mip_solve_rat_ri
Context.mip_solve_rat_batch = _mip_solve_rat_batch
Context.has_solution_batch = _has_solution_batch
