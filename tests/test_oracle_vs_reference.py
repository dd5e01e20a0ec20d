"""The oracle against the REAL reference library (oracle/_ref/libsgcref*.so compiled from
/root/reference): randomized cases beyond the committed fixtures.  CPU only; skipped where the
reference build is not available."""
import ctypes as C

import numpy as np
import pytest

import pyoracle as po


def _rand_box(rng, lo=-5, hi=5, ext=(2, 7)):
    l = rng.integers(lo, hi, 3)
    return l, l + rng.integers(ext[0], ext[1], 3) - 1


def test_reference_own_tests_passed():
    import os
    d = os.path.join(po.HERE, "_ref")
    if not os.path.isdir(d):
        pytest.skip("oracle/_ref not built")
    for t in ("testIntVect", "testBox", "testBoxIterator", "testBaseFab", "testDisjointBoxLayout",
              "testLayoutIterator", "testLevelData"):
        with open(os.path.join(d, t + ".log")) as f:
            assert "passed" in f.read(), t


def test_plan_random(oracle, reference):
    rng = np.random.default_rng(1)
    for _ in range(40):
        mb = rng.integers(2, 6, 3)
        nb = rng.integers(1, 5, 3)
        dlo = rng.integers(-4, 4, 3)
        dhi = dlo + mb * nb - 1
        ng = int(rng.integers(1, 3))
        if ng > mb.min():
            ng = 1
        per, trim = int(rng.integers(0, 8)), int(rng.integers(0, 8)) * 2
        nbox = int(np.prod(nb))
        nproc = int(rng.choice([p for p in (1, 2, 3, 4) if nbox % p == 0]))
        for p in range(nproc):
            a = oracle.copier_items(dlo, dhi, mb, ng, per, trim, nproc, p)
            b = reference.copier_items(dlo, dhi, mb, ng, 1, per, trim, nproc, p)
            assert np.array_equal(a[:, :24], b), (dlo, dhi, mb, ng, per, trim, nproc, p)


def test_send_region_is_mirror_of_recv(oracle):
    # m_regionSend (MPI-only in the reference, Copier.H:581-583) must equal the region the mirrored
    # item on the peer reads from: for item (A <- B, dir d) the peer item (B <- A, dir -d) reads
    # regionSendRemote == this item's m_regionSend.
    for per in (0, 7, 5):
        args = ((0, 0, 0), (11, 7, 15), (4, 4, 4), 1, per, 0)
        items = oracle.copier_items(*args)
        idx = {(int(r[0]), int(r[1]), tuple(r[14:17])): r for r in items}
        for r in items:
            m = idx[(int(r[1]), int(r[0]), tuple(-r[14:17]))]
            assert np.array_equal(r[24:30], m[8:14])


def test_exchange_random(oracle, reference):
    rng = np.random.default_rng(2)
    for _ in range(12):
        mb = rng.integers(3, 7, 3)
        nb = rng.integers(1, 4, 3)
        dlo = rng.integers(-4, 4, 3)
        dhi = dlo + mb * nb - 1
        ng, nc = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        per, trim = int(rng.integers(0, 8)), int(rng.integers(0, 8)) * 2
        boxes, owner, _ = oracle.layout(dlo, dhi, mb)
        items = oracle.copier_items(dlo, dhi, mb, ng, per, trim)
        u = rng.standard_normal(oracle.level_size(boxes, ng, nc))
        v = u.copy()
        oracle.level_exchange(boxes, ng, nc, items, u)
        n = reference.level_exchange(dlo, dhi, mb, ng, nc, per, trim, v, 0)
        assert n == len(items)
        assert np.array_equal(u, v)
        if per == 0:   # exchangeBegin/End: same-index copy, non-periodic only (LevelData.H:600)
            w = u.copy()
            v2 = rng.standard_normal(len(u))
            w2 = v2.copy()
            oracle.level_exchange(boxes, ng, nc, items, v2)
            reference.level_exchange(dlo, dhi, mb, ng, nc, per, trim, w2, 1)
            assert np.array_equal(v2, w2)


def test_heat_random_coefficients(oracle):
    rng = np.random.default_rng(3)
    refs = {1: po.Reference(""), 0: po.Reference("_nofma")} if po.have_reference("_nofma") else None
    if refs is None:
        pytest.skip("oracle/_ref not built")
    for contract, ref in refs.items():
        for f in (0.125, 1.0 / 6.0, 0.07, 0.3):
            dlo, dhi, mb = (0, 0, 0), (15, 11, 9), (4, 6, 5)
            boxes, _, _ = oracle.layout(dlo, dhi, mb)
            u = rng.standard_normal(oracle.level_size(boxes, 1, 1))
            v = u.copy()
            oracle.level_heat(dlo, dhi, mb, 1, 3, 0, f, 6, u, contract)
            ref.level_heat(dlo, dhi, mb, 1, 3, 0, f, 6, v)
            assert np.array_equal(u, v), (contract, f)


def test_fab_primitives_random(oracle, reference):
    rng = np.random.default_rng(4)
    L, R = oracle.lib, reference.lib
    i3 = po.i3
    for _ in range(30):
        dlo, dhi = _rand_box(rng, ext=(4, 9))
        slo, shi = _rand_box(rng, ext=(4, 9))
        dn, sn = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        ext = np.minimum(dhi - dlo, shi - slo) + 1
        rext = np.array([rng.integers(1, e + 1) for e in ext])
        rdlo = dlo + np.array([rng.integers(0, (dhi - dlo + 1)[d] - rext[d] + 1) for d in range(3)])
        rslo = slo + np.array([rng.integers(0, (shi - slo + 1)[d] - rext[d] + 1) for d in range(3)])
        nc = int(rng.integers(1, min(dn, sn) + 1))
        dc, sc = int(rng.integers(0, dn - nc + 1)), int(rng.integers(0, sn - nc + 1))
        flags = int(rng.integers(0, 16)) if rng.random() < 0.5 else 0xFFFFFFFF
        dsz = int(np.prod(dhi - dlo + 1)) * dn
        ssz = int(np.prod(shi - slo + 1)) * sn
        src = rng.standard_normal(ssz)
        d1 = rng.standard_normal(dsz)
        d2 = d1.copy()
        a = (i3(dlo), i3(dhi), dn)
        s = (i3(slo), i3(shi), sn)
        reg = (i3(rdlo), i3(rdlo + rext - 1), dc, i3(rslo), i3(rslo + rext - 1), sc, nc, flags)
        L.sgco_fab_copy(*a, d1, *s, src, *reg)
        R.sgcref_fab_copy(*a, d2, *s, src.copy(), *reg)
        assert np.array_equal(d1, d2)
        # linearOut / linearIn (BaseFab.cpp:350-419), wire order comp -> k -> j -> i
        c0 = int(rng.integers(0, dn))
        c1 = int(rng.integers(c0, dn + 1))
        n = int(np.prod(rext)) * dn
        b1, b2 = np.full(n, -7.0), np.full(n, -7.0)
        L.sgco_fab_linear_out(*a, d1, i3(rdlo), i3(rdlo + rext - 1), c0, c1, flags, b1)
        R.sgcref_fab_linear_out(*a, d2, i3(rdlo), i3(rdlo + rext - 1), c0, c1, flags, b2)
        assert np.array_equal(b1, b2)
        b1[:] = rng.standard_normal(n)
        L.sgco_fab_linear_in(*a, d1, i3(rdlo), i3(rdlo + rext - 1), c0, c1, flags, b1)
        R.sgcref_fab_linear_in(*a, d2, i3(rdlo), i3(rdlo + rext - 1), c0, c1, flags, b1.copy())
        assert np.array_equal(d1, d2)
        comp = int(rng.integers(-1, dn))
        L.sgco_fab_setval(*a, d1, comp, 3.25)
        R.sgcref_fab_setval(*a, d2, comp, 3.25)
        assert np.array_equal(d1, d2)


def test_adjbox_four_sign_cases(oracle, reference):
    # testBox.cpp:75-91
    for ncell in (-2, -1, 1, 3):
        for d in range(3):
            for side in (-1, 0, 1):
                a = oracle.adjbox((1, -2, 3), (6, 4, 9), ncell, d, side)
                b = reference.adjbox((1, -2, 3), (6, 4, 9), ncell, d, side)
                assert np.array_equal(a, b)


def test_laplacian_and_wave_random(oracle, reference):
    rng = np.random.default_rng(5)
    n = 20
    A = rng.standard_normal(n ** 3)
    B, _ = reference.laplacian(n, 16, A)
    assert np.array_equal(oracle.laplacian(n, 16, A), B)
    dlo, dhi = (0, 0, 0), (9, 7, 8)
    sz = 12 * 10 * 11
    u0, u1 = rng.standard_normal(sz), rng.standard_normal(sz)
    v0, v1 = u0.copy(), u1.copy()
    oracle.wave(dlo, dhi, 0.08, 5, u0, u1, 1)
    reference.wave(dlo, dhi, 0.08, 5, v0, v1)
    assert np.array_equal(u0, v0) and np.array_equal(u1, v1)


def test_lb_phases_and_steps_random(oracle, tmp_path, monkeypatch):
    """LBPatch::collision / macroscopic / stream one at a time, then whole LBLevel::advance steps, on
    random layouts (one and several box layers in z) against the reference LB application."""
    if not (po.have_reference("_lb") and po.have_reference("_lb_nofma")):
        pytest.skip("oracle/_ref LB build not available")
    monkeypatch.chdir(tmp_path)
    RN, RF = po.ReferenceLB("_nofma"), po.ReferenceLB("")
    rng = np.random.default_rng(11)
    for case in range(6):
        mb = rng.integers(2, 6, 3)
        nb = rng.integers(1, 4, 3)
        dlo = rng.integers(-3, 4, 3)
        dhi = dlo + mb * nb - 1
        boxes, owner, nbx = oracle.layout(dlo, dhi, mb)
        fi, mac = oracle.lb_initial(boxes)
        fi = fi * (1 + 0.05 * rng.standard_normal(fi.size))
        prev = fi * (1 + 0.05 * rng.standard_normal(fi.size))
        oracle.lb_macroscopic(boxes, fi, mac)
        # phases
        a = fi.copy(); oracle.lb_collision(boxes, a, mac)
        assert np.array_equal(a, RN.phase(dlo, dhi, mb, 0, fi, mac, prev)[0]), case
        m = np.zeros_like(mac); oracle.lb_macroscopic(boxes, fi, m)
        want = RN.phase(dlo, dhi, mb, 1, fi, np.zeros_like(mac), prev)[1]
        mask = want != 0                           # the reference never touches ghost cells of macro
        assert np.array_equal(m[mask], want[mask]) and np.array_equal(m, RF.phase(dlo, dhi, mb, 1, fi, np.zeros_like(mac), prev)[1])
        p = prev.copy(); oracle.lb_stream(boxes, fi, p)
        rfi, _, rprev = RN.phase(dlo, dhi, mb, 2, fi, mac, prev)      # the reference swaps: its m_curr is our p
        assert np.array_equal(p, rfi) and np.array_equal(fi, rprev), case
        # whole steps
        nstep = int(rng.integers(1, 6))
        f2, m2 = fi.copy(), mac.copy()
        p2 = oracle.lb_initial(boxes)[0]
        oracle.lb_advance(dlo, dhi, mb, nstep, f2, p2, m2)
        rfi, rmac, mass = RN.run(dlo, dhi, mb, nstep, fi.size, mac.size, fi, mac)
        assert np.array_equal(f2, rfi) and np.array_equal(m2, rmac), case
        ffi, fmac, _ = RF.run(dlo, dhi, mb, nstep, fi.size, mac.size, fi, mac)
        assert np.abs(f2 - ffi).max() <= 1e-12 * np.abs(ffi).max(), case
