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


# --------------------------------------------------------------------------
# std::mt19937_64 + std::uniform_real_distribution<double>(0, 1) as libstdc++ evaluates it
# (double(x) / 2^64, clamped below 1), vectorised over MANY independently seeded streams: the
# inputs SURVEY 8(d) names for the batched configurations (one seed per LP: 2024 + k at c2).
# Public algorithm (Matsumoto & Nishimura 2004).
# --------------------------------------------------------------------------
_NN, _MM = 312, 156
_MATRIX_A = np.uint64(0xB5026F5AA96619E9)
_UM, _LM = np.uint64(0xFFFFFFFF80000000), np.uint64(0x7FFFFFFF)


def _mt64_twist(mt):
    """One regeneration of the state array, all streams at once (mt: [B, 312] uint64, in place)."""
    one = np.uint64(1)

    def mix(hi, lo, far):
        x = (hi & _UM) | (lo & _LM)
        return far ^ (x >> one) ^ np.where((x & one) != 0, _MATRIX_A, np.uint64(0))
    # i in [0, 156): reads mt[i], mt[i+1] (not yet rewritten) and mt[i+156] (old)
    mt[:, :_MM] = mix(mt[:, :_MM], mt[:, 1:_MM + 1], mt[:, _MM:2 * _MM])
    # i in [156, 311): mt[i+1] still old, mt[i-156] already new
    mt[:, _MM:_NN - 1] = mix(mt[:, _MM:_NN - 1], mt[:, _MM + 1:_NN], mt[:, :_MM - 1])
    # i = 311: both neighbours already new
    mt[:, _NN - 1] = mix(mt[:, _NN - 1], mt[:, 0], mt[:, _MM - 1])


def mt64_uniform_many(seeds, count):
    """[len(seeds), count] float64: the first `count` draws of uniform_real_distribution(0, 1)
    from std::mt19937_64(seed), for every seed."""
    seeds = np.asarray(seeds, dtype=np.uint64)
    B = seeds.shape[0]
    out = np.empty((B, count), dtype=np.float64)
    with np.errstate(over="ignore"):
        mt = np.empty((B, _NN), dtype=np.uint64)
        mt[:, 0] = seeds
        for i in range(1, _NN):
            p = mt[:, i - 1]
            mt[:, i] = np.uint64(6364136223846793005) * (p ^ (p >> np.uint64(62))) + np.uint64(i)
        done = 0
        while done < count:
            _mt64_twist(mt)
            take = min(_NN, count - done)
            x = mt[:, :take].copy()
            x ^= (x >> np.uint64(29)) & np.uint64(0x5555555555555555)
            x ^= (x << np.uint64(17)) & np.uint64(0x71D67FFFEDA60000)
            x ^= (x << np.uint64(37)) & np.uint64(0xFFF7EEE000000000)
            x ^= x >> np.uint64(43)
            u = x.astype(np.float64) * (1.0 / 18446744073709551616.0)  # double(x): round to nearest, then exact scaling
            np.minimum(u, 0.99999999999999988897769753748, out=u)
            out[:, done:done + take] = u
            done += take
    return out


def dense_lp_batch(first_seed, batch, m, n, chunk=8192):
    """SURVEY 8(d) dense family for LPs first_seed .. first_seed + batch - 1, one std::mt19937_64
    stream per LP: per row all A_ij then b_i = 1 + U * n, then all c_j.  Returns leq [batch, m, n+1]
    and tgtf [batch, n+1]."""
    leq = np.empty((batch, m, n + 1), dtype=np.float64)
    tgtf = np.zeros((batch, n + 1), dtype=np.float64)
    for lo in range(0, batch, chunk):
        hi = min(batch, lo + chunk)
        u = mt64_uniform_many(np.arange(first_seed + lo, first_seed + hi, dtype=np.uint64), m * (n + 1) + n)
        leq[lo:hi] = u[:, :m * (n + 1)].reshape(hi - lo, m, n + 1)
        tgtf[lo:hi, :n] = u[:, m * (n + 1):]
    leq[:, :, n] = 1.0 + leq[:, :, n] * n
    return leq, tgtf
