"""CPU: the proof (tests/divconst_proof.py) that the LB collision kernel's fast division by its four
constants is correctly rounded; GPU: the device implementation against the device's IEEE division on
the enumerated near-boundary inputs, special values and random data."""
import random
from fractions import Fraction as Fr

import numpy as np
import pytest

import divconst_proof as dp


def test_fast_constant_division_is_correctly_rounded_for_all_inputs():
    for name, c in dp.CONSTANTS.items():
        cands, bad = dp.check_constant(name, c)
        assert not bad, (name, bad[:3])
        # and a much wider band of near-boundary inputs than the proof needs (|D| <= 300)
        wide, bad = dp.check_constant(name, c, 300)
        assert len(wide) >= 100 and not bad, (name, bad[:3])
    rnd = random.Random(1)
    for name, c in dp.CONSTANTS.items():
        rc = 1.0 / c
        for _ in range(5000):
            X = rnd.randrange(1 << 52, 1 << 53)
            q1, exact = dp.fast_div(X, c, rc)
            assert exact and q1 == float(Fr(X) / Fr(c)), (name, X)


@pytest.mark.gpu
def test_fast_constant_division_on_device(sgc):
    sgc.init(0)
    rng = np.random.default_rng(3)
    for which, (name, c) in enumerate(dp.CONSTANTS.items()):
        near, _ = dp.check_constant(name, c, 300)
        xs = [np.array(near, np.float64), -np.array(near, np.float64),
              np.array(near, np.float64) * 2.0 ** -60, np.array(near, np.float64) * 2.0 ** 40,
              rng.standard_normal(1 << 20), rng.standard_normal(1 << 18) * 1e-30, rng.standard_normal(1 << 18) * 1e30,
              np.array([0.0, -0.0, 5e-324, -5e-324, 1e-310, 2.0 ** -959, 2.0 ** -960, 2.0 ** 960, 2.0 ** 961, 1e308,
                        np.inf, -np.inf, np.nan, 1.0, c, 3 * c, c * c]),
              rng.integers(1 << 52, 1 << 53, 1 << 20).astype(np.float64)]
        x = np.concatenate(xs)
        fast, exact, cdev = sgc.lb_selftest_divconst(which, x)
        assert cdev == c
        assert np.array_equal(fast.view(np.uint64)[~np.isnan(exact)], exact.view(np.uint64)[~np.isnan(exact)]), name
        assert np.isnan(fast[np.isnan(exact)]).all()
        with np.errstate(all="ignore"):
            assert np.array_equal((x / c).view(np.uint64)[~np.isnan(exact)], exact.view(np.uint64)[~np.isnan(exact)])
