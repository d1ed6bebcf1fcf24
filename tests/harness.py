"""Shared test plumbing: ctypes bindings for the CPU checkers and LP generators.

TEST INFRASTRUCTURE ONLY.  `oracle()` is oracle/_build/libxp_oracle.so (our C
restatement, always buildable); `ref()` is oracle/_ref/libxpoly_ref.so (the
unmodified reference, present only where /root/reference was available at
build time).  Nothing under xpoly_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libxp_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libxpoly_ref.so")

SIX_SUCC, SIX_UNBOUND, SIX_NO_PRI, SIX_OPTINF, SIX_TIME_OUT = 0, 1, 2, 3, 4
IP_SUCC, IP_UNBOUND, IP_NO_PRI, IP_NO_BETTER = 0, 1, 2, 3
NO_LIMIT = 0xFFFFFFFF

_vp = C.c_void_p


def P(a):
    return None if a is None else a.ctypes.data_as(_vp)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        src_m = max(os.path.getmtime(os.path.join(ORACLE_DIR, f))
                    for f in ("xp_oracle.c", "xp_oracle.h", "xp_oracle_six.inc"))
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < src_m:
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.xo_appro_count.restype = C.c_longlong
        _oracle.xo_mt64_uniform.argtypes = [C.c_uint64, C.c_size_t, _vp]
        _oracle.xo_mt64_uniform.restype = None
    return _oracle


def ref():
    """The compiled reference, or None when oracle/_ref was not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = C.CDLL(REF_SO)
        _ref.ref_appro_count.restype = C.c_longlong
    return _ref


# --------------------------------------------------------------------------
# LP generators
# --------------------------------------------------------------------------
def mt64_uniform(seed, count):
    out = np.empty(count, dtype=np.float64)
    oracle().xo_mt64_uniform(seed, count, P(out))
    return out


def gen_dense_lp(seed, m, n):
    """SURVEY 8(d) family: A_ij~U(0,1), b_i = 1+U*n, c_j~U(0,1); draw order per
    row all A_ij then b_i, then all c_j (std::mt19937_64 stream)."""
    u = mt64_uniform(seed, m * (n + 1) + n)
    leq = u[: m * (n + 1)].reshape(m, n + 1).copy()
    leq[:, n] = 1.0 + leq[:, n] * n
    tgtf = np.zeros(n + 1)
    tgtf[:n] = u[m * (n + 1):]
    return leq, tgtf


def gen_mixed_lp(seed, m, n, lo=-1.0, hi=1.0, bneg=0.0):
    """Mixed-sign dense LP; bneg = probability of a negative rhs (forces phase 1)."""
    r = np.random.RandomState(seed)
    leq = r.uniform(lo, hi, size=(m, n + 1))
    leq[:, n] = 1.0 + r.uniform(0, 1, size=m) * n
    if bneg > 0:
        neg = r.uniform(0, 1, size=m) < bneg
        leq[neg, n] = -r.uniform(0.1, 2.0, size=int(neg.sum()))
    tgtf = np.zeros(n + 1)
    tgtf[:n] = r.uniform(lo, hi, size=n)
    return leq, tgtf


def gen_int_lp(seed, m, n, alo=0, ahi=3, density=0.3, blo=0, bhi=20, clo=1, chi=5):
    """SURVEY 8(d) c4 family: integer A at the given density, b, c.  Returned
    as float64 arrays holding integers (use to_rat() for the Rational side)."""
    r = np.random.RandomState(seed)
    A = r.randint(alo, ahi + 1, size=(m, n)).astype(np.float64)
    A *= (r.uniform(0, 1, size=(m, n)) < density)
    leq = np.zeros((m, n + 1))
    leq[:, :n] = A
    leq[:, n] = r.randint(blo, bhi + 1, size=m)
    tgtf = np.zeros(n + 1)
    tgtf[:n] = r.randint(clo, chi + 1, size=n)
    return leq, tgtf


def to_rat(a):
    """Integer-valued float/int array -> int32 array of (num, den) pairs."""
    a = np.asarray(a)
    out = np.empty(a.shape + (2,), dtype=np.int32)
    out[..., 0] = np.rint(a).astype(np.int32)
    out[..., 1] = 1
    return np.ascontiguousarray(out)


# --------------------------------------------------------------------------
# Thin wrappers (same shapes for the oracle `xo_*` and the reference `ref_*`)
# --------------------------------------------------------------------------
def _lib_and_prefix(which):
    if which == "oracle":
        return oracle(), "xo_"
    lib = ref()
    if lib is None:
        raise RuntimeError("oracle/_ref not built")
    return lib, "ref_"


def six_solve(which, kind, is_min, leq, tgtf, vc=None, eq=None, max_iter=NO_LIMIT):
    """kind: 'f64' (float64 arrays) or 'rat' (int32 (...,2) arrays)."""
    lib, pre = _lib_and_prefix(which)
    m, n1 = leq.shape[0], leq.shape[1]
    n = n1 - 1
    k = 0 if eq is None else eq.shape[0]
    if kind == "f64":
        v = np.zeros(1)
        sol = np.zeros(n1)
    else:
        v = np.zeros(2, dtype=np.int32)
        sol = np.zeros((n1, 2), dtype=np.int32)
    fn = getattr(lib, f"{pre}six_solve_{kind}")
    st = fn(int(is_min), m, n, P(np.ascontiguousarray(leq)), P(np.ascontiguousarray(tgtf)),
            P(None if vc is None else np.ascontiguousarray(vc)), k,
            P(None if eq is None else np.ascontiguousarray(eq)), C.c_uint32(max_iter), P(v), P(sol))
    return dict(status=st, v=v, sol=sol)


def two_stage(which, kind, leq, tgtf, max_iter=NO_LIMIT, want_log=False):
    lib, pre = _lib_and_prefix(which)
    m, n = leq.shape[0], leq.shape[1] - 1
    cap = n + m + 2
    dims = np.zeros(4, dtype=np.int32)
    if kind == "f64":
        tab = np.zeros(m * cap)
        otg = np.zeros(cap)
        maxv = np.zeros(1)
        ssol = np.zeros(cap)
    else:
        tab = np.zeros((m * cap, 2), dtype=np.int32)
        otg = np.zeros((cap, 2), dtype=np.int32)
        maxv = np.zeros(2, dtype=np.int32)
        ssol = np.zeros((cap, 2), dtype=np.int32)
    eq2bv = np.zeros(m, dtype=np.int32)
    bv2eq = np.zeros(cap, dtype=np.int32)
    nvset = np.zeros(cap, dtype=np.uint8)
    bvset = np.zeros(cap, dtype=np.uint8)
    args = [m, n, P(np.ascontiguousarray(leq)), P(np.ascontiguousarray(tgtf)),
            C.c_uint32(max_iter), P(dims), P(tab), P(otg), P(eq2bv), P(bv2eq), P(nvset),
            P(bvset), P(maxv), P(ssol)]
    log = None
    if which == "oracle":
        log_cap = 1 << 16 if want_log else 0
        logbuf = np.zeros((max(log_cap, 1), 4), dtype=np.int32)
        nlog = C.c_int(0)
        args += [P(logbuf), log_cap, C.byref(nlog)]
    st = getattr(lib, f"{pre}two_stage_{kind}")(*args)
    if which == "oracle" and want_log:
        log = logbuf[: min(nlog.value, log_cap)].copy()
    r, c = int(dims[0]), int(dims[1])
    tab2 = tab[: r * c].reshape((r, c) + tab.shape[1:]).copy()
    return dict(status=st, rows=r, cols=c, rhs_idx=int(dims[2]), sol_cols=int(dims[3]),
                tab=tab2, tgtf=otg[:c].copy(), eq2bv=eq2bv, bv2eq=bv2eq[:c - 1].copy(),
                nvset=nvset[:c - 1].copy(), bvset=bvset[:c - 1].copy(), maxv=maxv,
                slack_sol=ssol[: int(dims[3])].copy(), log=log)


def mip_solve_ri(which, is_min, is_bin, leq, tgtf, indicator, eq=None):
    """MIP<RMat,Rational>::maxm / minm with rational_indicator (n+1 flags)."""
    lib, pre = _lib_and_prefix(which)
    m, n1 = leq.shape[0], leq.shape[1]
    k = 0 if eq is None else eq.shape[0]
    v = np.zeros(2, dtype=np.int32)
    sol = np.zeros((n1, 2), dtype=np.int32)
    ind = np.ascontiguousarray(indicator, dtype=np.uint8)
    args = [int(is_min), int(is_bin), m, n1 - 1, P(np.ascontiguousarray(leq)), P(np.ascontiguousarray(tgtf)), k,
            P(None if eq is None else np.ascontiguousarray(eq)), P(ind), P(v), P(sol)]
    nodes = C.c_int(0)
    if which == "oracle":
        args.append(C.byref(nodes))
    st = getattr(lib, f"{pre}mip_solve_rat_ri")(*args)
    return dict(status=st, v=v, sol=sol, nodes=nodes.value)


def mip_solve(which, kind, is_min, is_bin, leq, tgtf, eq=None):
    lib, pre = _lib_and_prefix(which)
    m = 0 if leq is None else leq.shape[0]
    n1 = (leq if leq is not None else eq).shape[1]
    n = n1 - 1
    k = 0 if eq is None else eq.shape[0]
    if kind == "f64":
        v = np.zeros(1)
        sol = np.zeros(n1)
    else:
        v = np.zeros(2, dtype=np.int32)
        sol = np.zeros((n1, 2), dtype=np.int32)
    args = [int(is_min), int(is_bin), m, n, P(None if leq is None else np.ascontiguousarray(leq)),
            P(np.ascontiguousarray(tgtf)), k, P(None if eq is None else np.ascontiguousarray(eq)),
            P(v), P(sol)]
    nodes = C.c_int(0)
    if which == "oracle":
        args.append(C.byref(nodes))
    st = getattr(lib, f"{pre}mip_solve_{kind}")(*args)
    return dict(status=st, v=v, sol=sol, nodes=nodes.value)


def has_solution(which, leq, eq=None, is_int=True, is_unique=True):
    lib, pre = _lib_and_prefix(which)
    m = 0 if leq is None else leq.shape[0]
    k = 0 if eq is None else eq.shape[0]
    n = (leq if leq is not None else eq).shape[1] - 1
    return getattr(lib, f"{pre}has_solution_rat")(
        m, n, P(None if leq is None else np.ascontiguousarray(leq)), k,
        P(None if eq is None else np.ascontiguousarray(eq)), int(is_int), int(is_unique))


def appro_count(which):
    lib, pre = _lib_and_prefix(which)
    return getattr(lib, f"{pre}appro_count")()


def bits(a):
    """float64 array -> uint64 view for bit-exact comparison."""
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def slack_solve_oracle(kind, tab, tgtf, nvset, bvset, bv2eq, eq2bv, max_iter=NO_LIMIT,
                       vc_diag=None, vc_rhs=None, log_cap=1 << 16):
    """xo_slack_*: SIX::solveSlackForm alone on a caller-built slack form."""
    lib = oracle()
    m, Cc = tab.shape[0], tab.shape[1]
    if kind == "f64":
        tab = np.ascontiguousarray(tab, dtype=np.float64).copy()
        tgtf = np.ascontiguousarray(tgtf, dtype=np.float64).copy()
        maxv = np.zeros(1)
        sol = np.zeros(Cc)
    else:
        tab = np.ascontiguousarray(tab, dtype=np.int32).copy()
        tgtf = np.ascontiguousarray(tgtf, dtype=np.int32).copy()
        maxv = np.zeros(2, dtype=np.int32)
        sol = np.zeros((Cc, 2), dtype=np.int32)
    nvset = np.ascontiguousarray(nvset, dtype=np.uint8).copy()
    bvset = np.ascontiguousarray(bvset, dtype=np.uint8).copy()
    bv2eq = np.ascontiguousarray(bv2eq, dtype=np.int32).copy()
    eq2bv = np.ascontiguousarray(eq2bv, dtype=np.int32).copy()
    iters = np.zeros(1, dtype=np.uint32)
    log = np.zeros((max(log_cap, 1), 4), dtype=np.int32)
    nlog = C.c_int(0)
    st = getattr(lib, f"xo_slack_{kind}")(
        m, Cc, P(tab), P(tgtf), P(nvset), P(bvset), P(bv2eq), P(eq2bv), P(vc_diag), P(vc_rhs),
        C.c_uint32(max_iter), P(maxv), P(sol), P(iters), P(log), log_cap, C.byref(nlog))
    return dict(status=st, tab=tab, tgtf=tgtf, nvset=nvset, bvset=bvset, bv2eq=bv2eq,
                eq2bv=eq2bv, maxv=maxv, sol=sol, iters=int(iters[0]),
                log=log[: min(nlog.value, log_cap), 1:].copy())
