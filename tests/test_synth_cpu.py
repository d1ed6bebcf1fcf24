"""xpoly_b200.synth: the vectorised std::mt19937_64 streams behind the batched bench inputs
(one seed per LP) against the scalar C restatement the other generators use."""
import numpy as np

import harness as H
from xpoly_b200 import synth


def test_mt64_streams_match_the_scalar_generator():
    seeds = [0, 1, 5, 2024, 2025, 12345, 2 ** 63 + 7, 2 ** 64 - 1]
    u = synth.mt64_uniform_many(seeds, 1000)  # more than three regenerations of the state
    for k, s in enumerate(seeds):
        assert np.array_equal(H.bits(u[k]), H.bits(H.mt64_uniform(s, 1000))), s
    assert u.min() >= 0.0 and u.max() < 1.0


def test_dense_lp_batch_is_gen_dense_lp_per_seed():
    leq, tg = synth.dense_lp_batch(2024, 40, 32, 31, chunk=16)
    for k in (0, 1, 15, 16, 39):
        l, t = H.gen_dense_lp(2024 + k, 32, 31)
        assert np.array_equal(H.bits(leq[k]), H.bits(l)) and np.array_equal(H.bits(tg[k]), H.bits(t))
