"""The ALGORITHM of the two-steps-per-pass sweep (oracle/two_step_model.py: the data flow of
csrc/sgc_stream2.cu in numpy) against two { exchange ; sweep } steps of the oracle — on the CPU, so the
claim the kernel rests on is checked in the tier that runs without a GPU: every level-t+1 value a box needs,
those of its neighbours included, can be recomputed from valid cells (and, at faces without a motion item,
from the never-exchanged ghost cells) with the same operations on the same values, i.e. bit for bit."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import two_step_model as M  # noqa: E402


@pytest.mark.parametrize("dhi,mb,per", [
    ((15, 15, 11), (8, 8, 6), 7),        # 2 x 2 x 2 boxes, fully periodic
    ((7, 7, 5), (8, 8, 6), 7),           # one box, periodic onto itself
    ((23, 7, 11), (8, 8, 6), 7),         # 3 x 1 x 2
    ((15, 15, 11), (8, 8, 6), 0),        # physical boundaries all round
    ((15, 15, 11), (8, 8, 6), 3),        # z ends physical
    ((15, 23, 5), (8, 8, 6), 5),         # y ends physical, one box layer in z
    ((7, 15, 11), (8, 8, 6), 6),         # x ends physical, one box in x
    ((23, 15, 7), (8, 8, 4), 1),
    ((15, 15, 15), (8, 8, 8), 4),
])
def test_two_step_pass_equals_two_exchange_sweep_steps(oracle, dhi, mb, per):
    dlo = (0, 0, 0)
    boxes, _, _ = oracle.layout(dlo, dhi, mb)
    n = oracle.level_size(boxes, 1, 1)
    rng = np.random.default_rng(sum(dhi) + per)
    u0 = rng.standard_normal(n)                      # valid cells AND ghost cells (the static ones matter)
    f = 0.2
    want = u0.copy()
    oracle.level_heat(dlo, dhi, mb, 1, per, 0, f, 2, want, 0)
    items = oracle.copier_items(dlo, dhi, mb, 1, per, 0)
    nbr = M.face_neighbours(items)
    # the oracle ping-pongs against a copy of its input level, so both ghost levels are the input level
    fabs = M.split(u0, boxes)
    got = M.two_step_pass(fabs, fabs, fabs, nbr, f)
    wfabs = M.split(want, boxes)
    for b in range(len(boxes)):
        vz, V = fabs[b].shape[0] - 2, fabs[b].shape[1] - 2
        assert np.array_equal(got[b], wfabs[b][1:vz + 1, 1:V + 1, 1:V + 1]), (b, dhi, mb, per)


def test_two_step_pass_uses_each_levels_own_static_ghosts(oracle):
    """Faces without a motion item: step t reads the ghost cells of one level, step t+1 those of the other
    (the exchange never writes them).  Two explicit oracle steps with different ghost data in the two levels."""
    dlo, dhi, mb, per, f = (0, 0, 0), (15, 15, 11), (8, 8, 6), 2, 0.3
    boxes, _, _ = oracle.layout(dlo, dhi, mb)
    n = oracle.level_size(boxes, 1, 1)
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    items = oracle.copier_items(dlo, dhi, mb, 1, per, 0)
    fa, fb = M.split(a.copy(), boxes), M.split(b.copy(), boxes)
    got = M.two_step_pass(fa, fa, fb, M.face_neighbours(items), f)
    # reference: exchange(a); sweep a -> b; exchange(b); sweep b -> a
    oracle.level_exchange(boxes, 1, 1, items, a)
    oracle.level_heat_sweep(boxes, 1, a, b, f, 0)
    oracle.level_exchange(boxes, 1, 1, items, b)
    oracle.level_heat_sweep(boxes, 1, b, a, f, 0)
    wa = M.split(a, boxes)
    for k in range(len(boxes)):
        assert np.array_equal(got[k], wa[k][1:7, 1:9, 1:9]), k


def test_fma_form_of_the_update_is_exact():
    """csrc/sgc_stream2.cu forms the reference's (u[+e] - 2u) as ONE fma(-2, u, u[+e]) (heat_zhalf / heat_finish): 2u is
    exact in binary floating point, so both round the same real number once.  Checked on random doubles over the
    whole exponent range, on subnormals and on cancelling operands; the only inputs on which the two forms differ
    are those where 2u overflows (|u| >= 2^1023), which no heat field reaches."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(11)
    n = 20000
    mant = rng.uniform(1.0, 2.0, (2, n)) * rng.choice([-1.0, 1.0], (2, n))
    e_u = rng.integers(-1070, 1020, n)
    # the neighbour within a few binades of u (where cancellation and rounding matter) or anywhere
    e_p = np.where(rng.random(n) < 0.7, e_u + rng.integers(-3, 4, n), rng.integers(-1070, 1020, n))
    u = np.ldexp(mant[0], e_u)
    p = np.ldexp(mant[1], np.clip(e_p, -1074, 1023))
    cases = list(zip(u, p))
    cases += [(5e-324, 1e-323), (-5e-324, 5e-324), (2.2250738585072014e-308, 4.4501477170144023e-308), (1.0, 2.0), (1.0, 2.0000000000000004),
              (0.1, 0.2), (1e308 / 4, 1e308), (-0.0, 0.0), (0.0, -0.0), (3.0, 6.000000000000001)]
    for uu, pp in cases:
        ref = np.float64(pp) - np.float64(2.0) * np.float64(uu)          # the reference: multiply, then subtract
        got = libm.fma(-2.0, float(uu), float(pp))
        assert ref == got and np.signbit(ref) == np.signbit(got), (uu, pp, ref, got)
    # and the documented exception
    big = 1.5 * 2.0 ** 1023
    with np.errstate(over="ignore", invalid="ignore"):
        assert not np.isfinite(np.float64(big) - np.float64(2.0) * np.float64(big)) and np.isfinite(libm.fma(-2.0, big, big))
