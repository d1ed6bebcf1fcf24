"""Host mirror of the device-side synthetic LP generator (k_fill_synth in
csrc/xp_large_f64.cu): SURVEY 8(d) dense family, A_ij~U(0,1), b_i = 1+U*n,
c_j~U(0,1), from a counter-based splitmix64 stream, so the CPU checker and the
GPU see bit-identical inputs at any size without moving the matrix."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def u01(seed, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    h = _mix64(np.uint64(seed) ^ _mix64(idx))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def dense_lp(seed, m, n):
    """Normalised LP (leq m x (n+1), tgtf n+1) identical to fill_synthetic(seed)."""
    idx = np.arange(m * (n + 1), dtype=np.uint64).reshape(m, n + 1)
    leq = u01(seed, idx)
    leq[:, n] = 1.0 + leq[:, n] * n
    tgtf = np.zeros(n + 1)
    tgtf[:n] = u01(seed, np.uint64(m) * np.uint64(n + 1) + np.arange(n, dtype=np.uint64))
    return leq, tgtf
