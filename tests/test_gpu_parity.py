"""Parity tests proper: the CUDA path, called through the C ABI (libsgc_b200.so), against the
oracle and the golden vectors from the real reference.  Bit-exact for exchange / copy / pack /
index work; for the stencil bit-exact in both FMA modes (the reference's g++ -ffp-contract=fast
build and -ffp-contract=off), and <= 1e-12 relative across modes.

Restated reference tests: testLevelData.cpp:126-171 (exchange values on 2^3 boxes, 2 comps),
testBaseFab.cpp:142-257 (copy, shifted copy with comp remap, linearOut -> mutate -> linearIn),
testMPIExchange.cpp:90-153 (ghost face == neighbour id + 0.5), laplacian.cpp:80-86 (B == 0).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(sgc):
    sgc.init(0)
    return sgc


def t2_passes(nstep):
    """two-step passes of a single-rank sgc_heat_run on a covered level: (nstep - 2) // 2 — an odd count ends through
    the library's third level (a -> b -> scratch -> a), which needs at least three passes"""
    p = (nstep - 2) // 2
    return p if (p % 2 == 0 or p >= 3) else p & ~1


def make_level(g, dlo, dhi, mb, ncomp, ng, data=None, elem_bytes=8):
    dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
    lev = g.LevelData(dbl, ncomp, ng, elem_bytes)
    if data is not None:
        lev.upload(data)
    return dbl, lev


# ------------------------------------------------------------------ level container / transfers
def test_upload_download_roundtrip_and_setval(g, oracle):
    rng = np.random.default_rng(0)
    dbl, lev = make_level(g, (1, -1, 1), (12, 10, 12), (4, 4, 4), 2, 1)
    n = oracle.level_size(dbl.boxes, 1, 2)
    assert lev.elems == n
    assert [lev.fab_offset(b) for b in range(lev.size())] == list(oracle.level_offsets(dbl.boxes, 1, 2))
    u = rng.standard_normal(n)
    lev.upload(u)
    assert np.array_equal(lev.download(), u)
    # one fab at a time (BaseFab::copyToDevice / copyToHost)
    b = 5
    fab = rng.standard_normal(lev.fab_elems(b))
    lev.upload(fab, box=b)
    u[lev.fab_offset(b):lev.fab_offset(b) + fab.size] = fab
    assert np.array_equal(lev.download(), u)
    assert np.array_equal(lev.download(box=b), fab)
    # setVal(all) and setVal(comp) — testBaseFab / testLevelData.cpp:75-124
    lev.setVal(2.5)
    assert (lev.download() == 2.5).all()
    lev.setVal(-1.25, comp=1)
    got = lev.download()
    for k in range(lev.size()):
        f = got[lev.fab_offset(k):lev.fab_offset(k) + lev.fab_elems(k)].reshape(2, -1)
        assert (f[0] == 2.5).all() and (f[1] == -1.25).all()
    lev.setVal(7.0, comp=0, box=3)
    got = lev.download(box=3).reshape(2, -1)
    assert (got[0] == 7.0).all() and (got[1] == -1.25).all()


def test_valid_only_upload_and_download(g, oracle):
    """sgc_level_upload_valid / download_valid: the valid cells of every fab as one dense buffer
    [box][comp][k][j][i] (what LevelData.cpp:150-193 hands to the plot writer, and its reverse); ghost cells keep
    their contents on upload.  Blocking and transfer-stream forms, double and float, uniform and ragged boxes."""
    rng = np.random.default_rng(3)
    for dlo, dhi, mb, nc, dt in (((0, 0, 0), (127, 63, 63), (64, 64, 64), 1, np.float64),
                                 ((1, -1, 1), (12, 10, 12), (4, 4, 4), 2, np.float64),
                                 ((0, 0, 0), (31, 31, 15), (16, 16, 16), 1, np.float32)):
        dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
        lev = g.LevelData(dbl, nc, 1, dtype=dt)
        full = rng.standard_normal(lev.elems).astype(dt)
        lev.upload(full)
        nv = lev.validElems()
        assert nv == nc * int(np.prod(np.array(dhi) - np.array(dlo) + 1))
        vals = rng.standard_normal(nv).astype(dt)
        lev.uploadValid(vals)
        want = full.copy()
        pos = 0
        for k, bx in enumerate(dbl.boxes):
            n = bx[3:] - bx[:3] + 1
            f = want[lev.fab_offset(k):lev.fab_offset(k) + lev.fab_elems(k)].reshape(nc, n[2] + 2, n[1] + 2, n[0] + 2)
            cnt = nc * int(np.prod(n))
            f[:, 1:-1, 1:-1, 1:-1] = vals[pos:pos + cnt].reshape(nc, n[2], n[1], n[0])
            pos += cnt
        assert np.array_equal(lev.download(), want)              # valid cells replaced, ghost cells untouched
        assert np.array_equal(lev.downloadValid(), vals)
        # transfer-stream forms
        pin_in, pin_out = g.PinnedBuffer(nv, dt), g.PinnedBuffer(nv, dt)
        pin_in.array[:] = -vals
        st = g.Stream()
        lev.uploadValid(pin_in, st)
        st.sync()
        lev.downloadValid(pin_out, st)
        st.sync()
        assert np.array_equal(pin_out.array, -vals)


def test_async_transfer_streams_pipeline(g, oracle):
    """LevelData::copyToDeviceAsync / copyToHostAsync (LevelData.H:734-765): batches streamed through
    two transfer streams and the compute stream, ordered by events, give the same bits as the blocking
    calls — whatever the interleaving."""
    dlo, dhi, mb = (0, 0, 0), (31, 31, 31), (16, 16, 16)
    dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
    nset, nbatch, nstep = 2, 5, 6
    sets = []
    for _ in range(nset):
        la, lb = g.LevelData(dbl, 1, 1), g.LevelData(dbl, 1, 1)
        sets.append((la, lb, g.Copier().defineExchangeLD(la, 7, 0)))
    n = sets[0][0].elems
    rng = np.random.default_rng(5)
    ins = [g.PinnedBuffer(n) for _ in range(nbatch)]
    outs = [g.PinnedBuffer(n) for _ in range(nbatch)]
    for b in ins:
        b.array[:] = rng.integers(0, 1024, n) / 1024.0
    up, down = g.Stream(), g.Stream()
    ev_up = [g.Event() for _ in range(nset)]
    ev_run = [g.Event() for _ in range(nset)]
    ev_down = [g.Event() for _ in range(nset)]

    def submit_upload(j):
        k = j % nset
        if j >= nset:
            up.wait(ev_down[k])
        sets[k][0].uploadAsync(ins[j], up)
        ev_up[k].record(up)

    submit_upload(0)
    for j in range(nbatch):
        k = j % nset
        la, lb, cp = sets[k]
        if j + 1 < nbatch:
            submit_upload(j + 1)
        g.compute_wait(ev_up[k])
        lb.copyFrom(la)
        res = g.heat_run(la, lb, cp, 0.125, nstep)
        ev_run[k].record()
        down.wait(ev_run[k])
        res.downloadAsync(outs[j], down)
        ev_down[k].record(down)
    down.sync()
    for j in range(nbatch):
        want = ins[j].array.copy()
        oracle.level_heat(dlo, dhi, mb, 1, 7, 0, 0.125, nstep, want, contract=0)
        assert np.array_equal(outs[j].array, want), "batch %d" % j


def test_level_copy(g):
    rng = np.random.default_rng(12)
    dbl, a = make_level(g, (0, 0, 0), (15, 7, 7), (4, 4, 4), 2, 1)
    b = g.LevelData(dbl, 2, 1)
    u = rng.standard_normal(a.elems)
    a.upload(u)
    b.setVal(-3.0)
    b.copyFrom(a)
    assert np.array_equal(b.download(), u) and np.array_equal(a.download(), u)
    _, c = make_level(g, (0, 0, 0), (15, 7, 7), (4, 4, 4), 1, 1)
    with pytest.raises(g.SgcError):
        c.copyFrom(a)


def test_int_and_byte_fabs(g, oracle):
    # BaseFab<int> on an offset box (testBaseFab.cpp:114-140) and BaseFab<char>/<bool>
    rng = np.random.default_rng(1)
    for esz, dt in ((4, np.int32), (1, np.uint8)):
        dbl, lev = make_level(g, (-3, 2, 0), (4, 9, 7), (4, 4, 4), 2, 1, elem_bytes=esz)
        u = rng.integers(0, 100, lev.elems).astype(dt)
        lev.upload(u)
        cp = g.Copier().defineExchangeLD(lev, 7, 0)
        lev.exchange(cp)
        items = oracle.copier_items((-3, 2, 0), (4, 9, 7), (4, 4, 4), 1, 7, 0)
        want = u.astype(np.float64)
        oracle.level_exchange(dbl.boxes, 1, 2, items, want)
        assert np.array_equal(lev.download().astype(np.float64), want)


# ------------------------------------------------------------------ BaseFab::copy / linearOut / linearIn
def test_fab_copy_shifted_with_component_remap(g, oracle):
    # testBaseFab.cpp:169-209: shifted source box, comps 0..1 -> 1..2, sentinel elsewhere
    L = oracle.lib
    import pyoracle as po
    rng = np.random.default_rng(2)
    for _ in range(12):
        dlo = rng.integers(-4, 4, 3); dhi = dlo + rng.integers(4, 9, 3) - 1
        slo = rng.integers(-4, 4, 3); shi = slo + rng.integers(4, 9, 3) - 1
        dn, sn = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        rext = np.array([rng.integers(1, min(dhi[d] - dlo[d], shi[d] - slo[d]) + 2) for d in range(3)])
        rdlo = dlo + np.array([rng.integers(0, dhi[d] - dlo[d] + 2 - rext[d]) for d in range(3)])
        rslo = slo + np.array([rng.integers(0, shi[d] - slo[d] + 2 - rext[d]) for d in range(3)])
        nc = int(rng.integers(1, min(dn, sn) + 1))
        dc, sc = int(rng.integers(0, dn - nc + 1)), int(rng.integers(0, sn - nc + 1))
        flags = int(rng.integers(0, 8)) if rng.random() < 0.5 else 0xFFFFFFFF
        dst = g.LevelData(None, dn, 0, boxes=[list(dlo) + list(dhi)])
        src = g.LevelData(None, sn, 0, boxes=[list(slo) + list(shi)])
        d0, s0 = rng.standard_normal(dst.elems), rng.standard_normal(src.elems)
        dst.upload(d0); src.upload(s0)
        dst.copy(0, rdlo, rdlo + rext - 1, dc, src, 0, rslo, rslo + rext - 1, sc, nc, flags)
        want = d0.copy()
        L.sgco_fab_copy(po.i3(dlo), po.i3(dhi), dn, want, po.i3(slo), po.i3(shi), sn, s0,
                        po.i3(rdlo), po.i3(rdlo + rext - 1), dc, po.i3(rslo), po.i3(rslo + rext - 1), sc, nc, flags)
        assert np.array_equal(dst.download(), want)
        assert np.array_equal(src.download(), s0)


def test_linear_out_mutate_linear_in(g, oracle):
    # testBaseFab.cpp:212-257: wire order comp -> k -> j -> i
    L = oracle.lib
    import pyoracle as po
    rng = np.random.default_rng(3)
    lo, hi = np.array((-1, 0, 2)), np.array((5, 4, 7))
    fab = g.LevelData(None, 3, 0, boxes=[list(lo) + list(hi)])
    d0 = rng.standard_normal(fab.elems)
    fab.upload(d0)
    for (rlo, rhi, c0, c1, flags) in (((0, 1, 3), (3, 2, 6), 0, 3, 0xFFFFFFFF), ((-1, 0, 2), (5, 4, 7), 1, 2, 0xFFFFFFFF),
                                      ((2, 2, 2), (2, 2, 2), 0, 3, 0b101), ((0, 0, 3), (5, 0, 3), 0, 2, 0b10)):
        n = int(np.prod(np.array(rhi) - np.array(rlo) + 1)) * 3
        want = np.zeros(n)
        cnt = L.sgco_fab_linear_out(po.i3(lo), po.i3(hi), 3, d0, po.i3(rlo), po.i3(rhi), c0, c1, flags, want)
        got = fab.linearOut(0, rlo, rhi, c0, c1, flags)
        assert len(got) == cnt and np.array_equal(got, want[:cnt])
        got = got.copy()
        got[0] = -99.5
        fab.linearIn(0, rlo, rhi, c0, c1, got, flags)
        L.sgco_fab_linear_in(po.i3(lo), po.i3(hi), 3, d0, po.i3(rlo), po.i3(rhi), c0, c1, flags, got)
        assert np.array_equal(fab.download(), d0)


# ------------------------------------------------------------------ exchange
def test_exchange_neighbour_ids_testLevelData(g):
    # testLevelData.cpp:126-171: 2x2x2 boxes of 4^3, ng=1, 2 comps; box b holds +id / -id; after the
    # exchange every ghost region that lies in a neighbour holds that neighbour's id.
    dbl, lev = make_level(g, (0, 0, 0), (7, 7, 7), (4, 4, 4), 2, 1)
    u = np.zeros(lev.elems)
    for b in range(8):
        f = u[lev.fab_offset(b):lev.fab_offset(b) + lev.fab_elems(b)].reshape(2, -1)
        f[0], f[1] = b + 1, -(b + 1)
    lev.upload(u)
    cp = g.Copier().defineExchangeLD(lev)
    assert cp.numMotionItem() == 56 and cp.numRemoteItem() == 0
    lev.exchange(cp)
    got = lev.download()
    for b in range(8):
        f = got[lev.fab_offset(b):lev.fab_offset(b) + lev.fab_elems(b)].reshape(2, 6, 6, 6)   # comp, k, j, i
        lo = dbl.boxes[b][:3] - 1
        for k in range(6):
            for j in range(6):
                for i in range(6):
                    cell = lo + (i, j, k)
                    if ((cell < 0) | (cell > 7)).any():
                        want = b + 1          # outside the domain: untouched
                    else:
                        nb = (cell // 4)
                        want = int(nb[0] + 2 * nb[1] + 4 * nb[2]) + 1
                    assert f[0, k, j, i] == want and f[1, k, j, i] == -want


def test_exchange_matches_reference_golden(g, golden):
    meta, arr = golden
    for c in meta["exchange"]:
        dbl, lev = make_level(g, c["dlo"], c["dhi"], c["maxbox"], c["ncomp"], c["ng"], arr["xin_" + c["key"]])
        cp = g.Copier().defineExchangeLD(lev, c["periodic"], c["trim"])
        assert cp.numMotionItem() == c["nitem"]
        lev.exchange(cp)
        assert np.array_equal(lev.download(), arr["xout_" + c["key"]]), c["key"]
        if c["periodic"] == 0:
            # split-phase form (LevelData.H:577-677); serial reference: non-periodic only
            lev.upload(arr["xin_" + c["key"]])
            lev.exchangeBegin(cp)
            lev.exchangeEnd(cp)
            assert np.array_equal(lev.download(), arr["xout_" + c["key"]]), c["key"]


def test_exchange_random_vs_oracle(g, oracle):
    rng = np.random.default_rng(4)
    for _ in range(10):
        mb = rng.integers(3, 9, 3)
        nb = rng.integers(1, 4, 3)
        dlo = rng.integers(-4, 4, 3)
        dhi = dlo + mb * nb - 1
        ng, nc = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        per, trim = int(rng.integers(0, 8)), int(rng.integers(0, 8)) * 2
        dbl, lev = make_level(g, dlo, dhi, mb, nc, ng)
        u = rng.standard_normal(lev.elems)
        lev.upload(u)
        cp = g.Copier().defineExchangeLD(lev, per, trim)
        lev.exchange(cp)
        items = oracle.copier_items(dlo, dhi, mb, ng, per, trim)
        assert cp.numMotionItem() == len(items)
        oracle.level_exchange(dbl.boxes, ng, nc, items, u)
        assert np.array_equal(lev.download(), u), (dlo, dhi, mb, ng, nc, per, trim)


def test_exchange_component_range_and_flags(g, oracle):
    # Copier built for comps [1,3) of 4 (defineExchangeDBL startComp/numComp, Copier.H:521-533) and a
    # per-item compRecvFlags mask (Motion2Way::setCompRecvFlags, Copier.H:430-436)
    rng = np.random.default_rng(5)
    dlo, dhi, mb = (0, 0, 0), (7, 7, 3), (4, 4, 4)
    dbl, lev = make_level(g, dlo, dhi, mb, 4, 1)
    u = rng.standard_normal(lev.elems)
    lev.upload(u)
    cp = g.Copier().defineExchangeLD(lev, 7, 0, comp_begin=1, ncomp=2)
    lev.exchange(cp)
    items = oracle.copier_items(dlo, dhi, mb, 1, 7, 0)
    want = u.copy()
    oracle.level_exchange(dbl.boxes, 1, 4, items, want, c0=1, nc=2)
    assert np.array_equal(lev.download(), want)
    # masked: only comp 2 of the range moves for odd items
    it = dbl.copier_items(1, 7, 0)
    it["comp_flags"][1::2] = 0b0100
    items[1::2, 23] = 0b0100
    lev.upload(u)
    cp2 = g.Copier().defineItems(lev, it, comp_begin=1, ncomp=2)
    lev.exchange(cp2)
    want = u.copy()
    oracle.level_exchange(dbl.boxes, 1, 4, items, want, c0=1, nc=2)
    assert np.array_equal(lev.download(), want)
    # the same edited masks through the exchange plan itself (sgc_plan_create_exchange_ex: what the facade's
    # Copier does after setCompRecvFlags, and the form that also carries remote items)
    lev.upload(u)
    cp3 = g.Copier().defineExchangeLD(lev, 7, 0, comp_begin=1, ncomp=2, item_flags=it["comp_flags"])
    lev.exchange(cp3)
    assert np.array_equal(lev.download(), want)
    with pytest.raises(g.SgcError):
        g.Copier().defineExchangeLD(lev, 7, 0, comp_begin=1, ncomp=2, item_flags=it["comp_flags"][:-1])


def test_copier_narrower_than_the_levels_ghost_width(g, oracle):
    """Copier::defineExchangeDBL(dbl, numGhost, ...) (Copier.H:521-528) with numGhost smaller than the level's
    ghost width: only the inner ghost layer is filled; and one plan per level geometry (a Copier serves every
    LevelData on its layout)."""
    rng = np.random.default_rng(15)
    dlo, dhi, mb = (0, 0, 0), (11, 7, 7), (4, 4, 4)
    dbl, lev = make_level(g, dlo, dhi, mb, 2, 2)
    u = rng.standard_normal(lev.elems)
    lev.upload(u)
    cp = g.Copier().defineExchangeLD(lev, 5, 0, nghost=1)
    lev.exchange(cp)
    want = u.copy()
    oracle.level_exchange(dbl.boxes, 2, 2, oracle.copier_items(dlo, dhi, mb, 1, 5, 0), want)
    assert np.array_equal(lev.download(), want)
    full = g.Copier().defineExchangeLD(lev, 5, 0)
    lev.exchange(full)
    oracle.level_exchange(dbl.boxes, 2, 2, oracle.copier_items(dlo, dhi, mb, 2, 5, 0), want)
    assert np.array_equal(lev.download(), want)
    with pytest.raises(g.SgcError):
        g.Copier().defineExchangeLD(lev, 5, 0, nghost=3)
    _, other = make_level(g, dlo, dhi, mb, 1, 2)
    assert cp.matches(lev) and not cp.matches(other)


def test_plan_geometry_mismatch_is_an_error(g):
    _, a = make_level(g, (0, 0, 0), (7, 7, 7), (4, 4, 4), 1, 1)
    _, b = make_level(g, (0, 0, 0), (7, 7, 7), (4, 4, 4), 2, 1)
    cp = g.Copier().defineExchangeLD(a)
    with pytest.raises(g.SgcError):
        b.exchange(cp)                      # different layout tag (LevelData.H:285)
    with pytest.raises(g.SgcError):
        g.heat_sweep(a, a, 0.1)


# ------------------------------------------------------------------ stencil
def test_heat_matches_reference_golden(g, golden, oracle):
    meta, arr = golden
    for c in meta["heat"]:
        dbl, a = make_level(g, c["dlo"], c["dhi"], c["maxbox"], 1, 1)
        b = g.LevelData(dbl, 1, 1)
        u0 = oracle.fill_hash(dbl.boxes, 1, 0.0)
        a.upload(u0); b.upload(u0)
        cp = g.Copier().defineExchangeLD(a, c["periodic"], 0)
        res = g.heat_run(a, b, cp, c["f"], c["nstep"], g.SGC_FMA if c["contract"] else 0)
        assert np.array_equal(res.download(), arr["heat_" + c["key"]]), c["key"]


def test_heat_known_answer_hashes(g, golden, oracle):
    # 64^3 periodic, hash init, 100 sweeps: f=0.125 -> 978934eb8172d8e9 for box sizes
    # 16/32/64 in either FMA mode; f=1/6 -> fb0472d7beb7e3a6 (contracted) / 1d09ec6f6796b0be (not)
    meta, _ = golden
    for c in meta["heat_hash"]:
        dbl, a = make_level(g, c["dlo"], c["dhi"], c["maxbox"], 1, 1)
        b = g.LevelData(dbl, 1, 1)
        u0 = oracle.fill_hash(dbl.boxes, 1, 0.0)
        modes = (1, 0) if c["f"] == 0.125 else (c["contract"],)
        for contract in modes:
            a.upload(u0); b.upload(u0)
            cp = g.Copier().defineExchangeLD(a, c["periodic"], 0)
            assert cp.numMotionItem() == c["nitem"]
            res = g.heat_run(a, b, cp, c["f"], c["nstep"], g.SGC_FMA if contract else 0)
            out = res.download()
            assert "%016x" % oracle.fnv1a64(out) == c["fnv_level"], (c["key"], contract)
            glob = oracle.gather(c["dlo"], c["dhi"], dbl.boxes, 1, 1, 0, out)
            assert "%016x" % oracle.fnv1a64(glob) == c["fnv"], (c["key"], contract)


def test_heat_random_vs_oracle_both_fma_modes(g, oracle):
    rng = np.random.default_rng(6)
    for (dlo, dhi, mb, per) in (((0, 0, 0), (15, 11, 9), (4, 6, 5), 3), ((-3, 1, 2), (28, 16, 9), (8, 4, 8), 7),
                                ((0, 0, 0), (69, 5, 3), (70, 3, 2), 0)):
        for f in (0.125, 1.0 / 6.0, 0.31):
            for contract in (1, 0):
                dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
                b = g.LevelData(dbl, 1, 1)
                u0 = rng.standard_normal(a.elems)
                a.upload(u0); b.upload(u0)
                cp = g.Copier().defineExchangeLD(a, per, 0)
                res = g.heat_run(a, b, cp, f, 5, g.SGC_FMA if contract else 0)
                want = u0.copy()
                oracle.level_heat(dlo, dhi, mb, 1, per, 0, f, 5, want, contract)
                assert np.array_equal(res.download(), want), (dlo, dhi, mb, f, contract)


def test_two_step_sweep_matches_oracle_and_step_by_step(g, oracle, monkeypatch):
    """sgc_heat_run's two-steps-per-pass kernel (sgc_stream2.cu; 64x64 and 32x32 boxes, periodic or with
    physical boundaries): level t is read once and level t+2 written once.  Valid AND ghost cells of BOTH levels
    must equal the oracle's nstep x {exchange; sweep} (LevelData.H:455-566 + the sweep loop) and the
    library's own step-by-step schedule, bit for bit, in both FMA modes; odd and even step counts,
    several boxes per direction, flat boxes (vz = 8).  (Re-boxing of the 32^3-box levels into 64^3 unions is
    switched off here: this test is about the kernel instances themselves.)"""
    monkeypatch.setenv("SGC_NO_REBOX", "1")
    rng = np.random.default_rng(21)
    cases = [((0, 0, 0), (63, 63, 63), (64, 64, 64), 7, 0.125, 10, 1),       # one box, periodic onto itself
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 7, 0.31, 7, 1),
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 7, 1.0 / 6.0, 12, 0),
             ((0, 0, 0), (191, 63, 127), (64, 64, 64), 7, 0.2, 9, 1),
             ((-64, 0, 8), (-1, 127, 23), (64, 64, 8), 7, 0.2, 11, 0),
             # physical boundaries (faces without a motion item keep their ghost cells, each level its own):
             # none / some / all directions periodic, one and several boxes per direction
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 0, 0.2, 9, 1),
             ((0, 0, 0), (191, 127, 63), (64, 64, 64), 0, 1.0 / 6.0, 8, 0),
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 3, 0.125, 10, 1),
             ((0, 0, 0), (191, 127, 127), (64, 64, 64), 5, 0.3, 7, 1),
             ((0, 0, 0), (127, 191, 15), (64, 64, 8), 6, 0.2, 11, 0),
             ((0, 0, 0), (63, 63, 63), (64, 64, 64), 1, 0.2, 8, 1),
             ((0, 0, 0), (127, 127, 63), (64, 64, 64), 2, 0.2, 8, 1),
             ((0, 0, 0), (63, 127, 127), (64, 64, 64), 4, 0.2, 13, 1),
             # the 32 x 32 instance (8 warps, 4 rows per thread, three CTAs per SM)
             ((0, 0, 0), (31, 31, 31), (32, 32, 32), 7, 0.2, 10, 1),
             ((0, 0, 0), (95, 63, 63), (32, 32, 32), 7, 1.0 / 6.0, 9, 0),
             ((-32, 0, 4), (31, 95, 19), (32, 32, 8), 7, 0.3, 12, 1),
             ((0, 0, 0), (63, 63, 63), (32, 32, 32), 0, 0.2, 8, 0),
             ((0, 0, 0), (95, 63, 31), (32, 32, 16), 5, 0.3, 7, 1),
             ((0, 0, 0), (63, 95, 63), (32, 32, 32), 2, 0.3, 11, 1)]
    for (dlo, dhi, mb, per, f, nstep, contract) in cases:
        dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
        b = g.LevelData(dbl, 1, 1)
        u0 = rng.standard_normal(a.elems)
        v0 = rng.standard_normal(a.elems)          # the second level starts with different ghosts
        cp = g.Copier().defineExchangeLD(a, per, 0)
        flags = g.SGC_FMA if contract else 0
        outs = []
        for no_t2 in (False, True):
            if no_t2:
                monkeypatch.setenv("SGC_NO_T2", "1")
            else:
                monkeypatch.delenv("SGC_NO_T2", raising=False)
            a.upload(u0); b.upload(v0)
            g.profile_enable(True)
            g.profile_read(2)
            res = g.heat_run(a, b, cp, f, nstep, flags)
            n2, _ = g.profile_read(2)
            g.profile_enable(False)
            assert (n2 == 0) == no_t2, "two-step passes launched: %d (SGC_NO_T2 %s)" % (n2, no_t2)
            assert n2 in (0, t2_passes(nstep))
            assert (res is b) == bool(nstep & 1)
            outs.append((a.download(), b.download()))
        monkeypatch.delenv("SGC_NO_T2", raising=False)
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), (dhi, mb, per, "vs step-by-step")
        if per != 7:
            # the oracle ping-pongs against a copy of ITS input level, ghosts included: give it equal ghosts
            a.upload(u0); b.upload(u0)
            outs[0] = None
            res = g.heat_run(a, b, cp, f, nstep, flags)
            outs[0] = (a.download(), b.download())
        want = u0.copy()
        oracle.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want, contract)
        got = outs[0][1] if nstep & 1 else outs[0][0]
        # the oracle ping-pongs against a copy of the input level; ghost cells of the result level were
        # last written by the exchange two steps earlier, which both schedules reproduce — compare the
        # valid cells against the oracle and everything (above) against the step-by-step schedule
        assert np.array_equal(oracle.gather(dlo, dhi, dbl.boxes, 1, 1, 0, got),
                              oracle.gather(dlo, dhi, dbl.boxes, 1, 1, 0, want)), (dhi, mb, per, "vs oracle")


def test_two_step_sweep_marches_down_pillars(g, oracle, monkeypatch):
    """The two-step kernel marches down PILLARS of z-stacked boxes (sgc_plan.cpp: pillar_structure; halo planes two
    boxes share are read and swept once).  The plan reports the structure; with the pillars switched off
    (SGC_T2_NO_PILLARS: box by box, the round-1 traversal) BOTH levels come out with the same bits, ghost cells
    included; and the valid cells equal the oracle's step-by-step run.  Pillars of 3 boxes whose top box wraps onto
    their own bottom box, 2 x 2 pillars of 2, flat boxes in pillars of 4, odd and even pass counts."""
    cases = [((0, 0, 0), (63, 63, 191), (64, 64, 64), 7, 0.2, 9, (1, 3)),
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 7, 0.125, 12, (4, 2)),
             ((0, 0, 0), (63, 127, 31), (64, 64, 8), 7, 0.3, 11, (2, 4)),
             ((0, 0, 0), (127, 63, 127), (64, 64, 64), 3, 0.2, 8, (2, 2))]     # physical z ends: the boundary instance marches box by box
    rng = np.random.default_rng(31)
    for dlo, dhi, mb, per, f, nstep, want_pillars in cases:
        dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
        b = g.LevelData(dbl, 1, 1)
        u0 = rng.standard_normal(a.elems)
        cp = g.Copier().defineExchangeLD(a, per, 0)
        assert cp.pillars() == want_pillars, (dhi, mb, cp.pillars())
        outs = []
        for off in (False, True):
            if off:
                monkeypatch.setenv("SGC_T2_NO_PILLARS", "1")
            else:
                monkeypatch.delenv("SGC_T2_NO_PILLARS", raising=False)
            a.upload(u0); b.upload(u0)
            g.profile_enable(True)
            g.profile_read(2)
            res = g.heat_run(a, b, cp, f, nstep)
            n2, _ = g.profile_read(2)
            g.profile_enable(False)
            assert n2 == t2_passes(nstep) and (res is b) == bool(nstep & 1)
            outs.append((a.download(), b.download()))
        monkeypatch.delenv("SGC_T2_NO_PILLARS", raising=False)
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), (dhi, mb, "pillars vs box by box")
        want = u0.copy()
        oracle.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want, 1)
        got = outs[0][1] if nstep & 1 else outs[0][0]
        assert np.array_equal(oracle.gather(dlo, dhi, dbl.boxes, 1, 1, 0, got),
                              oracle.gather(dlo, dhi, dbl.boxes, 1, 1, 0, want)), (dhi, mb, per, "vs oracle")


def test_two_step_sweep_physical_boundary_and_fallback(g, oracle):
    """Non-periodic z: boxes at the domain boundary have no z neighbour; the two-step kernel's boundary
    instance reads the (never exchanged) ghost planes instead — every cell of the result, ghost cells
    included, equals the oracle's step-by-step run."""
    dlo, dhi, mb = (0, 0, 0), (127, 127, 127), (64, 64, 64)
    dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
    b = g.LevelData(dbl, 1, 1)
    u0 = np.random.default_rng(22).standard_normal(a.elems)
    a.upload(u0); b.upload(u0)
    cp = g.Copier().defineExchangeLD(a, 3, 0)
    g.profile_enable(True)
    g.profile_read(2)
    res = g.heat_run(a, b, cp, 0.125, 8)
    n2, _ = g.profile_read(2)
    g.profile_enable(False)
    assert n2 == t2_passes(8) == 3
    want = u0.copy()
    oracle.level_heat(dlo, dhi, mb, 1, 3, 0, 0.125, 8, want, 1)
    assert np.array_equal(res.download(), want)
    # boxes with an odd number of planes are not covered either (CTA ranges start on even planes)
    dlo, dhi, mb = (0, 0, 0), (63, 63, 13), (64, 64, 7)
    dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
    b = g.LevelData(dbl, 1, 1)
    u0 = np.random.default_rng(24).standard_normal(a.elems)
    a.upload(u0); b.upload(u0)
    cp = g.Copier().defineExchangeLD(a, 7, 0)
    g.profile_enable(True)
    g.profile_read(2)
    res = g.heat_run(a, b, cp, 0.2, 9)
    n2, _ = g.profile_read(2)
    g.profile_enable(False)
    assert n2 == 0
    want = u0.copy()
    oracle.level_heat(dlo, dhi, mb, 1, 7, 0, 0.2, 9, want, 1)
    assert np.array_equal(res.download(), want)


def test_exchange_after_sweep_reads_edge_strips(g, oracle, monkeypatch):
    """The drop-in path of an unmodified application — LevelData::exchange (LevelData.H:455-566), then the
    sweep, step after step.  The streaming sweep leaves dense copies of its first / last x column (edge
    strips) and the next exchange takes its x faces, x edges and corners from them; every cell of both
    levels, ghost cells included, must equal the oracle's exchange + sweep after every step, and any other
    writer of valid cells (upload, setVal, BaseFab::copy, linearIn) must switch the exchange back to the
    fabs themselves."""
    rng = np.random.default_rng(31)
    for (dlo, dhi, mb, per, trim) in (((0, 0, 0), (127, 127, 63), (64, 64, 32), 7, 0),
                                      ((0, 0, 0), (127, 63, 63), (64, 64, 64), 0, 0),
                                      ((-32, 0, 0), (63, 63, 31), (32, 32, 16), 5, 0),
                                      ((0, 0, 0), (95, 63, 31), (32, 32, 32), 7, 8),          # TrimCorner
                                      ((0, 0, 0), (127, 127, 7), (64, 64, 4), 3, 4)):         # TrimEdge
        dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
        b = g.LevelData(dbl, 1, 1)
        items = oracle.copier_items(dlo, dhi, mb, 1, per, trim)
        hu, hv = rng.standard_normal(a.elems), rng.standard_normal(a.elems)
        a.upload(hu); b.upload(hv)
        assert a.edgeStrips() == 0
        cp = g.Copier().defineExchangeLD(a, per, trim)
        cur, nxt, hc, hn = a, b, hu, hv
        for step in range(4):
            cur.exchange(cp)
            oracle.level_exchange(dbl.boxes, 1, 1, items, hc)
            assert np.array_equal(cur.download(), hc), (mb, per, step, "exchange")
            g.heat_sweep(nxt, cur, 0.2)
            oracle.level_heat_sweep(dbl.boxes, 1, hc, hn, 0.2, 1)
            assert nxt.edgeStrips() == 2, "the whole-level streaming sweep keeps the edge strips current"
            assert np.array_equal(nxt.download(), hn), (mb, per, step, "sweep")
            cur, nxt, hc, hn = nxt, cur, hn, hc
        # cur's strips are current; every other writer of valid cells must invalidate them
        n = (mb[0] + 2) * (mb[1] + 2) * (mb[2] + 2)
        fab = rng.standard_normal(n)
        cur.upload(fab, box=1)
        hc[cur.fab_offset(1):cur.fab_offset(1) + n] = fab
        assert cur.edgeStrips() == 0
        cur.exchange(cp)
        oracle.level_exchange(dbl.boxes, 1, 1, items, hc)
        assert np.array_equal(cur.download(), hc), (mb, per, "after upload")
        for writer in ("setval", "copy", "linear_in"):
            g.heat_sweep(nxt, cur, 0.125)
            oracle.level_heat_sweep(dbl.boxes, 1, hc, hn, 0.125, 1)
            cur, nxt, hc, hn = nxt, cur, hn, hc
            assert cur.edgeStrips() == 2
            lo = [int(x) for x in dbl.boxes[0][:3]]
            f0 = hc[:n].reshape(mb[2] + 2, mb[1] + 2, mb[0] + 2)
            if writer == "setval":
                cur.setVal(-3.0, comp=0, box=0)
                f0[...] = -3.0
            elif writer == "copy":            # first valid x column of box 0 <- its second one
                r0, r1 = (lo[0], lo[1], lo[2]), (lo[0], lo[1] + mb[1] - 1, lo[2] + mb[2] - 1)
                s0, s1 = (lo[0] + 1, lo[1], lo[2]), (lo[0] + 1, lo[1] + mb[1] - 1, lo[2] + mb[2] - 1)
                cur.copy(0, r0, r1, 0, cur, 0, s0, s1, 0, 1)
                f0[1:-1, 1:-1, 1] = f0[1:-1, 1:-1, 2].copy()
            else:
                r0, r1 = (lo[0], lo[1], lo[2]), (lo[0], lo[1] + 1, lo[2] + 1)
                cur.linearIn(0, r0, r1, 0, 1, np.arange(4.0))
                f0[1:3, 1:3, 1] = np.arange(4.0).reshape(2, 2)
            assert cur.edgeStrips() == 0, writer
            cur.exchange(cp)
            oracle.level_exchange(dbl.boxes, 1, 1, items, hc)
            assert np.array_equal(cur.download(), hc), (mb, per, writer)
        # the strip-sourced table and the plain one are interchangeable (SGC_NO_XE selects the plain one)
        g.heat_sweep(nxt, cur, 0.125)
        oracle.level_heat_sweep(dbl.boxes, 1, hc, hn, 0.125, 1)
        monkeypatch.setenv("SGC_NO_XE", "1")
        nxt.exchange(cp)
        monkeypatch.delenv("SGC_NO_XE")
        oracle.level_exchange(dbl.boxes, 1, 1, items, hn)
        assert np.array_equal(nxt.download(), hn)


def test_reboxed_step_loop_small_boxes(g, oracle, monkeypatch):
    """Levels of many small boxes (BASELINE: 32 768 boxes of 16^3; the reference's own multi-box application
    uses 16^3 boxes, latticeBoltzmann.cpp:16,36): sgc_heat_run advances them on 64^2-footprint unions of the
    user's boxes with the two-step sweep and hands the result back.  Valid AND ghost cells of BOTH levels
    must equal the library's own per-box schedule (SGC_NO_REBOX) bit for bit, the valid cells the oracle's
    nstep x {exchange; sweep}; periodic and physical boundaries, both FMA modes, odd and even step counts."""
    rng = np.random.default_rng(33)
    cases = [((0, 0, 0), (127, 127, 127), (16, 16, 16), 7, 0.125, 13, 1, (4, 4, 4)),
             ((0, 0, 0), (63, 127, 63), (16, 16, 16), 0, 0.2, 10, 0, (4, 4, 4)),
             ((0, 0, 0), (127, 63, 95), (16, 16, 16), 5, 1.0 / 6.0, 9, 1, (4, 4, 3)),
             ((0, 0, 0), (127, 63, 63), (32, 32, 32), 3, 0.3, 12, 1, (2, 2, 2)),
             ((-64, 0, 0), (-1, 63, 31), (8, 8, 8), 6, 0.2, 8, 1, (8, 8, 4)),
             ((0, 0, 0), (63, 63, 47), (16, 32, 12), 7, 0.2, 11, 0, (4, 2, 4))]
    for (dlo, dhi, mb, per, f, nstep, contract, want_c) in cases:
        dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
        b = g.LevelData(dbl, 1, 1)
        u0, v0 = rng.standard_normal(a.elems), rng.standard_normal(a.elems)
        flags = g.SGC_FMA if contract else 0
        outs = []
        for no_rebox in (False, True):
            if no_rebox:
                monkeypatch.setenv("SGC_NO_REBOX", "1")
            else:
                monkeypatch.delenv("SGC_NO_REBOX", raising=False)
            cp = g.Copier().defineExchangeLD(a, per, 0)
            a.upload(u0); b.upload(v0)
            g.profile_enable(True)
            g.profile_read(2)
            res = g.heat_run(a, b, cp, f, nstep, flags)
            n2, _ = g.profile_read(2)
            g.profile_enable(False)
            decided, fac = cp.reboxFactors()
            assert decided and fac == ((1, 1, 1) if no_rebox else want_c), (mb, fac)
            if not no_rebox:
                assert n2 == ((nstep - 2) & ~1) // 2, "two-step passes on the union fabs"
            assert (res is b) == bool(nstep & 1)
            outs.append((a.download(), b.download()))
        monkeypatch.delenv("SGC_NO_REBOX", raising=False)
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), (dhi, mb, per, "vs per-box schedule")
        # oracle: ping-pongs against a copy of its input level -> give both levels the same ghosts
        cp = g.Copier().defineExchangeLD(a, per, 0)
        a.upload(u0); b.upload(u0)
        res = g.heat_run(a, b, cp, f, nstep, flags)
        want = u0.copy()
        oracle.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want, contract)
        assert np.array_equal(res.download(), want), (dhi, mb, per, "vs oracle")


def test_heat_tolerance_across_fma_modes(g, oracle):
    # the north_star's floating-point bar: <= 1e-12 relative per cell after N steps (observed CPU
    # fast-vs-off gap after 100 steps: 1.6e-15)
    dlo, dhi, mb = (0, 0, 0), (31, 31, 31), (16, 16, 16)
    dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
    b = g.LevelData(dbl, 1, 1)
    u0 = oracle.fill_hash(dbl.boxes, 1, 0.0)
    outs = []
    for flags in (g.SGC_FMA, 0):
        a.upload(u0); b.upload(u0)
        cp = g.Copier().defineExchangeLD(a, 7, 0)
        outs.append(g.heat_run(a, b, cp, 1.0 / 6.0, 100, flags).download())
    rel = np.abs(outs[0] - outs[1]) / np.maximum(np.abs(outs[1]), 1e-300)
    assert rel.max() <= 1e-12


def test_single_sweep_leaves_ghosts_and_input_untouched(g, oracle):
    rng = np.random.default_rng(8)
    dlo, dhi, mb = (0, 0, 0), (11, 7, 9), (6, 4, 5)
    dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
    b = g.LevelData(dbl, 1, 1)
    u, v = rng.standard_normal(a.elems), rng.standard_normal(a.elems)
    a.upload(u); b.upload(v)
    g.heat_sweep(b, a, 0.2)
    want = v.copy()
    oracle.level_heat_sweep(dbl.boxes, 1, u, want, 0.2, 1)
    assert np.array_equal(b.download(), want)       # ghosts of unew keep their old contents
    assert np.array_equal(a.download(), u)
    # box sub-range (used for interior/boundary overlap)
    b.upload(v)
    g.heat_sweep(b, a, 0.2, box_range=(2, 5))
    got = b.download()
    for k in range(a.size()):
        s = slice(a.fab_offset(k), a.fab_offset(k) + a.fab_elems(k))
        assert np.array_equal(got[s], want[s] if 2 <= k < 5 else v[s])


def test_single_precision_and_multicomponent_sweeps(g, oracle):
    """Real = float (the reference's USE_SINGLE_PRECISION build, Parameters.H:17-23,197): nstep x {exchange; sweep}
    on float levels against the reference library itself compiled that way (oracle/_ref/libsgcref_sp*.so),
    both FMA modes; and multi-component levels: every component is swept (double, against the oracle per
    component)."""
    import pyoracle
    rng = np.random.default_rng(41)
    for variant, flags in (("_sp", g.SGC_FMA), ("_sp_nofma", 0)):
        if not pyoracle.have_reference(variant):
            pytest.skip("oracle/_ref/libsgcref%s.so not built" % variant)
        R = pyoracle.Reference(variant)
        for (dlo, dhi, mb, per, f, nstep) in (((0, 0, 0), (23, 15, 15), (8, 8, 8), 7, 0.2, 6),
                                              ((0, 0, 0), (127, 63, 63), (64, 64, 64), 3, 1.0 / 6.0, 5),
                                              ((0, 0, 0), (63, 31, 31), (32, 32, 32), 7, 0.2, 4),
                                              ((0, 0, 0), (31, 31, 15), (16, 16, 16), 5, 0.3, 5),
                                              ((0, 0, 0), (11, 9, 8), (6, 5, 3), 0, 0.31, 4)):
            dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
            a, b = g.LevelData(dbl, 1, 1, dtype=np.float32), g.LevelData(dbl, 1, 1, dtype=np.float32)
            u0 = rng.standard_normal(a.elems).astype(np.float32)
            a.upload(u0); b.upload(u0)
            cp = g.Copier().defineExchangeLD(a, per, 0)
            res = g.heat_run(a, b, cp, f, nstep, flags)
            want = u0.copy()
            R.level_heat(dlo, dhi, mb, 1, per, 0, np.float32(f), nstep, want)
            got = res.download()
            assert got.dtype == np.float32 and np.array_equal(got, want), (variant, mb)
    # double and float levels do not mix
    dbl = g.DisjointBoxLayout(g.Box((0, 0, 0), (7, 7, 7)), 4)
    with pytest.raises(g.SgcError):
        g.heat_sweep(g.LevelData(dbl, 1, 1), g.LevelData(dbl, 1, 1, dtype=np.float32), 0.1)
    # three components, double: each is an independent field (table-driven kernel, and the streaming kernel,
    # which walks component c of fab b as virtual fab 3 b + c)
    for dlo, dhi, mb in (((0, 0, 0), (15, 11, 7), (8, 4, 4)), ((0, 0, 0), (63, 31, 15), (32, 32, 8)),
                         ((0, 0, 0), (63, 127, 11), (64, 64, 6))):
        dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
        a, b = g.LevelData(dbl, 3, 1), g.LevelData(dbl, 3, 1)
        u, v = rng.standard_normal(a.elems), rng.standard_normal(a.elems)
        a.upload(u); b.upload(v)
        g.heat_sweep(b, a, 0.2)
        got = b.download()
        n = [int(np.prod(bx[3:] - bx[:3] + 3)) for bx in dbl.boxes]
        for c in range(3):
            uc = np.concatenate([u[a.fab_offset(k) + c * n[k]:a.fab_offset(k) + (c + 1) * n[k]] for k in range(a.size())])
            vc = np.concatenate([v[a.fab_offset(k) + c * n[k]:a.fab_offset(k) + (c + 1) * n[k]] for k in range(a.size())])
            gc = np.concatenate([got[a.fab_offset(k) + c * n[k]:a.fab_offset(k) + (c + 1) * n[k]] for k in range(a.size())])
            oracle.level_heat_sweep(dbl.boxes, 1, uc, vc, 0.2, 1)
            assert np.array_equal(gc, vc), (mb, c)


def test_two_ghost_layers_streaming_and_step_loop(g, oracle):
    """Levels with nghost = 2 (Copier plans exist for any ghost width, Copier.H:521-528): the 7-point sweep reads the
    inner layer only — streaming instances for 64^2 / 32^2 / 16^2 boxes, the table-driven kernel otherwise — and the
    step loop exchanges both layers every step.  Against the oracle, every cell of both levels."""
    rng = np.random.default_rng(47)
    for dlo, dhi, mb, per in (((0, 0, 0), (127, 63, 15), (64, 64, 8), 7), ((0, 0, 0), (63, 31, 31), (32, 32, 16), 3),
                              ((-16, 0, 0), (15, 31, 15), (16, 16, 16), 5), ((0, 0, 0), (11, 9, 8), (6, 5, 3), 0)):
        dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
        a, b = g.LevelData(dbl, 1, 2), g.LevelData(dbl, 1, 2)
        u, v = rng.standard_normal(a.elems), rng.standard_normal(a.elems)
        a.upload(u); b.upload(v)
        g.heat_sweep(b, a, 0.2)
        want = v.copy()
        oracle.level_heat_sweep(dbl.boxes, 2, u, want, 0.2, 1)
        assert np.array_equal(b.download(), want), (mb, "sweep")
        assert np.array_equal(a.download(), u)
        a.upload(u); b.upload(u)
        cp = g.Copier().defineExchangeLD(a, per, 0)
        res = g.heat_run(a, b, cp, 0.125, 5, 0)
        want = u.copy()
        oracle.level_heat(dlo, dhi, mb, 2, per, 0, 0.125, 5, want, 0)
        assert np.array_equal(res.download(), want), (mb, "run")


def test_reference_device_kats_gpuSandbox(g):
    """The reference's device-side known-answer tests restated through this library (gpuSandbox.cpp:66-151;
    its own kernels live in the application's sandbox_Cuda.cu on CudaFab / SlabFab): fab on grow((-1..2)^3, 1)
    with 2 components — test 3 set every value, tests 4/5 set the core cells only, test 6 the two-component
    3-point z stencil B = A[-e_z] + A + A[+e_z] on the core box, everything else of B untouched."""
    core_lo, core_hi = (-1, -1, -1), (2, 2, 2)
    A = g.LevelData(None, 2, 1, boxes=[list(core_lo) + list(core_hi)])
    n = 6
    # test 3: all values of comp 0 / 1
    A.setVal(-1.0, comp=0); A.setVal(-2.0, comp=1)
    f = A.download().reshape(2, n, n, n)
    assert (f[0] == -1.0).all() and (f[1] == -2.0).all()
    # tests 4 and 5: core cells only (BaseFab::copy(box, src) from a constant fab, BaseFab.cpp:254-271)
    for vals in ((1.0, 2.0), (3.0, 4.0)):
        K = g.LevelData(None, 2, 1, boxes=[list(core_lo) + list(core_hi)])
        K.setVal(vals[0], comp=0); K.setVal(vals[1], comp=1)
        A.copy(0, core_lo, core_hi, 0, K, 0, core_lo, core_hi, 0, 2)
        f = A.download().reshape(2, n, n, n)
        core = np.zeros((n, n, n), bool); core[1:-1, 1:-1, 1:-1] = True
        assert (f[0][core] == vals[0]).all() and (f[1][core] == vals[1]).all()
        assert (f[0][~core] == -1.0).all() and (f[1][~core] == -2.0).all()
    # test 6: A[c0] = i0, A[c1] = i0 + 1 on the ghost box, B preset to 1.234567E89
    i0 = np.arange(-2, 4, dtype=np.float64)[None, None, :] + np.zeros((n, n, n))
    hostA = np.stack([i0, i0 + 1.0])
    A.upload(hostA.ravel())
    B = g.LevelData(None, 2, 1, boxes=[list(core_lo) + list(core_hi)])
    B.setVal(1.234567e89)
    g.stencil7_apply(B, A, (1, 0, 0, 0, 0, 1, 1))
    fb = B.download().reshape(2, n, n, n)
    for c in range(2):
        want = hostA[c][:-2, 1:-1, 1:-1] + hostA[c][1:-1, 1:-1, 1:-1] + hostA[c][2:, 1:-1, 1:-1]
        assert np.array_equal(fb[c][1:-1, 1:-1, 1:-1], want)
        ghost = np.ones((n, n, n), bool); ghost[1:-1, 1:-1, 1:-1] = False
        assert (fb[c][ghost] == 1.234567e89).all()
    # and a general weighted stencil on random data, double and float, without contraction (numpy has no fma)
    rng = np.random.default_rng(43)
    w = rng.standard_normal(7)
    for dt in (np.float64, np.float32):
        dbl = g.DisjointBoxLayout(g.Box((0, 0, 0), (9, 7, 5)), (5, 4, 3))
        a, b = g.LevelData(dbl, 2, 1, dtype=dt), g.LevelData(dbl, 2, 1, dtype=dt)
        u = rng.standard_normal(a.elems).astype(dt)
        a.upload(u); b.setVal(0.0)
        g.stencil7_apply(b, a, w, flags=0)
        got = b.download()
        wt = w.astype(dt)
        for k, bx in enumerate(dbl.boxes):
            m = bx[3:] - bx[:3] + 3
            fu = u[a.fab_offset(k):a.fab_offset(k) + a.fab_elems(k)].reshape(2, m[2], m[1], m[0])
            fg = got[a.fab_offset(k):a.fab_offset(k) + a.fab_elems(k)].reshape(2, m[2], m[1], m[0])
            c = fu[:, 1:-1, 1:-1, 1:-1]
            r = wt[0] * c
            for wk, nb in zip(wt[1:], (fu[:, 1:-1, 1:-1, :-2], fu[:, 1:-1, 1:-1, 2:], fu[:, 1:-1, :-2, 1:-1],
                                        fu[:, 1:-1, 2:, 1:-1], fu[:, :-2, 1:-1, 1:-1], fu[:, 2:, 1:-1, 1:-1])):
                r = r + wk * nb
            assert r.dtype == dt and np.array_equal(fg[:, 1:-1, 1:-1, 1:-1], r), dt


def test_laplacian_driver_C1(g, golden, oracle):
    # application/laplacian/laplacian.cpp: A on (1..n)^3, B on (2..n-1)^3, 16 iterations
    meta, arr = golden
    for c in meta["laplacian"]:
        n = c["n"]
        A = g.LevelData(None, 1, 0, boxes=[[1, 1, 1, n, n, n]])
        B = g.LevelData(None, 1, 0, boxes=[[2, 2, 2, n - 1, n - 1, n - 1]])
        if c["init"] == "lin":
            ii = np.arange(1, n + 1, dtype=np.float64)
            a0 = (ii[None, None, :] + ii[None, :, None] + ii[:, None, None]).ravel().copy()
        else:
            a0 = oracle.fill_hash(np.array([[1, 1, 1, n, n, n]], np.int32), 0)
        A.upload(a0)
        for fuse in (0, 1):
            B.setVal(0.0)
            g.laplacian_apply(B, A, 0.5, c["niter"], fuse)
            got = B.download()
            assert np.array_equal(got, arr[c["key"]]), (c["key"], fuse)
            if c["init"] == "lin":
                assert not got.any()        # laplacian.cpp:80-86,115-120
    # the BASELINE config-1 size: single 128^3 box, linear field -> exactly zero; hash field vs oracle
    n = 128
    A = g.LevelData(None, 1, 0, boxes=[[1, 1, 1, n, n, n]])
    B = g.LevelData(None, 1, 0, boxes=[[2, 2, 2, n - 1, n - 1, n - 1]])
    ii = np.arange(1, n + 1, dtype=np.float64)
    A.upload((ii[None, None, :] + ii[None, :, None] + ii[:, None, None]).ravel().copy())
    B.setVal(0.0)
    g.laplacian_apply(B, A, 0.5, 16, 0)
    assert not B.download().any()
    a0 = oracle.fill_hash(np.array([[1, 1, 1, n, n, n]], np.int32), 0)
    A.upload(a0)
    B.setVal(0.0)
    g.laplacian_apply(B, A, 0.5, 16, 0)
    assert np.array_equal(B.download(), oracle.laplacian(n, 16, a0))


def test_wave_and_zero_gradient_bc(g, golden, oracle):
    # WavePatch::advance (WavePatch.cpp:136-254): BC face copies + leap-frog update, scalar form
    meta, arr = golden
    for c in meta["wave"]:
        dom = g.Box(c["dlo"], c["dhi"])
        box = [list(c["dlo"]) + list(c["dhi"])]
        lv = [g.LevelData(None, 1, 1, boxes=box) for _ in range(3)]
        u0 = oracle.fill_hash(np.array(box, np.int32), 1, 0.0)
        lv[0].upload(u0); lv[2].upload(u0); lv[1].setVal(0.0)
        bc = [g.Copier().defineBCZeroGradient(l, dom) for l in lv]
        n, np1, nm1 = 0, 1, 2
        for _ in range(c["nstep"]):
            lv[n].exchange(bc[n])
            g.wave_step(lv[np1], lv[n], lv[nm1], c["factor"], g.SGC_FMA if c["contract"] else 0)
            n, np1, nm1 = np1, nm1, n
        assert np.array_equal(lv[n].download(), arr[c["key"] + "_u0"]), c["key"]
        assert np.array_equal(lv[nm1].download(), arr[c["key"] + "_u1"]), c["key"]
    # multi-box BC plan: faces of a 2x2x1 layout
    rng = np.random.default_rng(9)
    dlo, dhi = (0, 0, 0), (7, 7, 3)
    dbl, lev = make_level(g, dlo, dhi, (4, 4, 4), 1, 1)
    u = rng.standard_normal(lev.elems)
    lev.upload(u)
    lev.exchange(g.Copier().defineBCZeroGradient(lev, g.Box(dlo, dhi)))
    got = lev.download()
    for b in range(4):
        f = u[lev.fab_offset(b):lev.fab_offset(b) + lev.fab_elems(b)].reshape(6, 6, 6).copy()
        bl, bh = dbl.boxes[b][:3], dbl.boxes[b][3:]
        if bl[0] == 0: f[1:5, 1:5, 0] = f[1:5, 1:5, 1]
        if bh[0] == 7: f[1:5, 1:5, 5] = f[1:5, 1:5, 4]
        if bl[1] == 0: f[1:5, 0, 1:5] = f[1:5, 1, 1:5]
        if bh[1] == 7: f[1:5, 5, 1:5] = f[1:5, 4, 1:5]
        f[0, 1:5, 1:5] = f[1, 1:5, 1:5]
        f[5, 1:5, 1:5] = f[4, 1:5, 1:5]
        assert np.array_equal(got[lev.fab_offset(b):lev.fab_offset(b) + lev.fab_elems(b)], f.ravel()), b


def test_shipped_wave_app_form(g, golden):
    # the reference's shipped time stepper (WavePatch::advance, AVX pencil path): BC + update in the
    # MD_DIRSUM form, valid cells bit-identical to the golden vectors produced by the real WavePatch
    meta, arr = golden
    for c in meta["wavepatch"]:
        h = c["h"]
        dom = g.Box((0, 0, 0), (h - 1,) * 3)
        box = [[0, 0, 0, h - 1, h - 1, h - 1]]
        lv = [g.LevelData(None, 1, 1, boxes=box) for _ in range(3)]
        lv[0].upload(arr[c["key"] + "_u_init"]); lv[2].upload(arr[c["key"] + "_um1_init"]); lv[1].setVal(0.0)
        bc = [g.Copier().defineBCZeroGradient(l, dom) for l in lv]
        n, np1, nm1 = 0, 1, 2
        for _ in range(c["nstep"]):
            lv[n].exchange(bc[n])
            g.wave_step(lv[np1], lv[n], lv[nm1], c["factor"], g.SGC_FMA | g.SGC_WAVE_DIRSUM)
            n, np1, nm1 = np1, nm1, n
        m = h + 2
        for lev, key in ((lv[n], "_un"), (lv[nm1], "_unm1")):
            got = lev.download().reshape(m, m, m)[1:-1, 1:-1, 1:-1]
            want = arr[c["key"] + key].reshape(m, m, m)[1:-1, 1:-1, 1:-1]
            assert np.array_equal(got, want), (c["key"], key)


def test_wave_run_device_resident_loop(g, golden):
    # sgc_wave_run = nstep x {BC; update; rotate} on the device, against the real WavePatch golden vectors
    meta, arr = golden
    for c in meta["wavepatch"]:
        h = c["h"]
        dom = g.Box((0, 0, 0), (h - 1,) * 3)
        box = [[0, 0, 0, h - 1, h - 1, h - 1]]
        lv = [g.LevelData(None, 1, 1, boxes=box) for _ in range(3)]
        lv[0].upload(arr[c["key"] + "_u_init"]); lv[2].upload(arr[c["key"] + "_um1_init"]); lv[1].setVal(0.0)
        bc = g.Copier().defineBCZeroGradient(lv[0], dom)     # same geometry: one BC plan serves all three levels
        un, _, unm1 = g.wave_run(lv, None, bc, c["factor"], c["nstep"], g.SGC_FMA | g.SGC_WAVE_DIRSUM)
        m = h + 2
        for lev, key in ((un, "_un"), (unm1, "_unm1")):
            got = lev.download().reshape(m, m, m)[1:-1, 1:-1, 1:-1]
            assert np.array_equal(got, arr[c["key"] + key].reshape(m, m, m)[1:-1, 1:-1, 1:-1]), (c["key"], key)


def test_wave_level_streaming_and_generic(g, oracle):
    # leap-frog update on box-decomposed levels (streaming instances 64/32/16 and the generic kernel),
    # both FMA modes, against the oracle's per-box restatement of WavePatch.cpp:226-244
    rng = np.random.default_rng(13)
    for mb, nb in (((64, 64, 64), (1, 1, 2)), ((32, 32, 32), (2, 1, 2)), ((16, 16, 16), (2, 2, 3)), ((6, 4, 5), (2, 2, 2))):
        dlo = (0, 0, 0)
        dhi = tuple(mb[d] * nb[d] - 1 for d in range(3))
        dbl, un = make_level(g, dlo, dhi, mb, 1, 1)
        unm1, unp1 = g.LevelData(dbl, 1, 1), g.LevelData(dbl, 1, 1)
        a, b, c = (rng.standard_normal(un.elems) for _ in range(3))
        cp = g.Copier().defineExchangeLD(un, 7, 0)
        items = oracle.copier_items(dlo, dhi, mb, 1, 7, 0)
        for contract in (1, 0):
            un.upload(a); unm1.upload(b); unp1.upload(c)
            un.exchange(cp)
            g.wave_step(unp1, un, unm1, 0.07, g.SGC_FMA if contract else 0)
            wa = a.copy()
            oracle.level_exchange(dbl.boxes, 1, 1, items, wa)
            want = c.copy()
            oracle.level_wave_sweep(dbl.boxes, 1, wa, b, want, 0.07, contract)
            assert np.array_equal(unp1.download(), want), (mb, contract)
            assert np.array_equal(unm1.download(), b)


def test_sweep_box_ranges_streaming_kernel(g, oracle):
    # the streaming kernel on sub-ranges and on two ranges in one launch (boundary box layers)
    rng = np.random.default_rng(11)
    for mb, nbz in ((16, 6), (64, 3), (32, 4), (8, 5)):
        dlo, dhi = (0, 0, 0), (2 * mb - 1, mb - 1, nbz * mb - 1)
        dbl, a = make_level(g, dlo, dhi, (mb,) * 3, 1, 1)
        b = g.LevelData(dbl, 1, 1)
        u, v = rng.standard_normal(a.elems), rng.standard_normal(a.elems)
        a.upload(u)
        want = v.copy()
        oracle.level_heat_sweep(dbl.boxes, 1, u, want, 0.3, 1)
        n = a.size()
        for rngs in ((0, n), (1, n - 1), (0, 2, n - 2, n), (0, 1, 3, 4), (2, 2, 3, n), (0, 3, 3, 3)):
            b.upload(v)
            g.heat_sweep(b, a, 0.3, box_range=rngs)
            got = b.download()
            sel = set(range(rngs[0], rngs[1])) | (set(range(rngs[2], rngs[3])) if len(rngs) == 4 else set())
            for k in range(n):
                s = slice(a.fab_offset(k), a.fab_offset(k) + a.fab_elems(k))
                assert np.array_equal(got[s], want[s] if k in sel else v[s]), (mb, rngs, k)


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_C2_properties(g, oracle):
    """BASELINE configs[1] shape (512^3 in 64^3 boxes is 2.4 GB for two levels; run 256^3 here and the
    full size in bench.py): (i) exchange fills every ghost with the periodic image of the global
    field; (ii) the sweep is linear: S(a*u + v) == a*S(u) + S(v) exactly for dyadic data;
    (iii) a constant field is a fixed point; (iv) the global sum is conserved to rounding."""
    n, mb = 256, 64
    dlo, dhi = (0, 0, 0), (n - 1,) * 3
    dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
    b = g.LevelData(dbl, 1, 1)
    cp = g.Copier().defineExchangeLD(a, 7, 0)
    assert cp.numMotionItem() == 26 * 64 and cp.numCells() == 64 * (66 ** 3 - 64 ** 3)
    # (i) ghost == periodic image: use u = global linear index (exact in double)
    glob = np.arange(n ** 3, dtype=np.float64).reshape(n, n, n)          # [k, j, i]
    host = np.full(a.elems, -1.0)
    for k in range(a.size()):
        lo = dbl.boxes[k][:3]
        f = host[a.fab_offset(k):a.fab_offset(k) + a.fab_elems(k)].reshape(66, 66, 66)
        f[1:65, 1:65, 1:65] = glob[lo[2]:lo[2] + 64, lo[1]:lo[1] + 64, lo[0]:lo[0] + 64]
    a.upload(host)
    a.exchange(cp)
    got = a.download()
    idx = [np.arange(-1, 65)] * 3
    for k in (0, 21, 63):
        lo = dbl.boxes[k][:3]
        f = got[a.fab_offset(k):a.fab_offset(k) + a.fab_elems(k)].reshape(66, 66, 66)
        kk, jj, ii = np.meshgrid((idx[2] + lo[2]) % n, (idx[1] + lo[1]) % n, (idx[0] + lo[0]) % n, indexing="ij")
        assert np.array_equal(f, glob[kk, jj, ii])
    # (iii) constant field is a fixed point of exchange + sweep
    a.setVal(3.5); b.setVal(0.0)
    res = g.heat_run(a, b, cp, 0.125, 3)
    out = res.download()
    for k in (0, 37):
        f = out[a.fab_offset(k):a.fab_offset(k) + a.fab_elems(k)].reshape(66, 66, 66)
        assert (f[1:65, 1:65, 1:65] == 3.5).all()
    # (ii) linearity on dyadic data and (iv) conservation
    u = oracle.fill_hash(dbl.boxes, 1, 0.0)
    rng = np.random.default_rng(10)
    v = rng.integers(0, 256, a.elems) / 256.0

    def sweep_valid(x):
        a.upload(x); b.setVal(0.0)
        a.exchange(cp)
        g.heat_sweep(b, a, 0.125)
        return b.download()
    su, sv, suv = sweep_valid(u), sweep_valid(v), sweep_valid(4.0 * u + v)
    assert np.array_equal(suv, 4.0 * su + sv)
    mask = np.zeros((66, 66, 66), bool)
    mask[1:65, 1:65, 1:65] = True
    m = np.tile(mask.ravel(), a.size())
    assert abs(su[m].sum() - u[m].sum()) <= 1e-9 * abs(u[m].sum())
    # (v) the same two properties through the step loop, whose passes advance TWO steps at a time on this
    # level (sgc_stream2.cu): 8 steps of dyadic data with f = 1/8 stay exact in double, so the run is linear
    # bit for bit, valid cells and ghost cells

    def run8(x):
        a.upload(x); b.upload(x)
        g.profile_enable(True); g.profile_read(2)
        r = g.heat_run(a, b, cp, 0.125, 8)
        n2, _ = g.profile_read(2); g.profile_enable(False)
        assert n2 == t2_passes(8) == 3 and r is a
        return r.download()
    ru, rv, ruv = run8(u), run8(v), run8(4.0 * u + v)
    assert np.array_equal(ruv, 4.0 * ru + rv)
    assert abs(ru[m].sum() - u[m].sum()) <= 1e-9 * abs(u[m].sum())


def test_two_step_sweep_256cubed_vs_oracle(g, oracle):
    """A quarter-size C2 level (256^3 in 64^3 boxes, 64 boxes, every box with 26 distinct neighbours or
    periodic images) through 10 steps = 4 two-step passes + fused + plain steps, against the oracle's
    step-by-step run: every cell of the result level, ghost cells included."""
    n, mb = 256, 64
    dlo, dhi = (0, 0, 0), (n - 1,) * 3
    dbl, a = make_level(g, dlo, dhi, mb, 1, 1)
    b = g.LevelData(dbl, 1, 1)
    cp = g.Copier().defineExchangeLD(a, 7, 0)
    u0 = np.random.default_rng(23).standard_normal(a.elems)
    a.upload(u0); b.upload(u0)
    res = g.heat_run(a, b, cp, 1.0 / 6.0, 10, g.SGC_FMA)
    want = u0.copy()
    oracle.level_heat(dlo, dhi, (mb,) * 3, 1, 7, 0, 1.0 / 6.0, 10, want, 1)
    assert np.array_equal(res.download(), want)


# ------------------------------------------------------------------ copy sets + lattice Boltzmann application
def test_copyset_replays_fab_copies_in_one_launch(g, oracle):
    """A list of BaseFab::copy calls (BaseFab.cpp:296-329) between two multi-box levels with different
    component counts, shifts into ghost cells and flags == the oracle's sequential loop."""
    import pyoracle as po
    L = oracle.lib
    rng = np.random.default_rng(21)
    dlo, dhi, mb = (0, -2, 1), (11, 5, 8), (4, 4, 4)
    dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
    dst, src = g.LevelData(dbl, 3, 1), g.LevelData(dbl, 5, 1)
    d0, s0 = rng.standard_normal(dst.elems), rng.standard_normal(src.elems)
    dst.upload(d0); src.upload(s0)
    ops = np.zeros(dbl.size() * 2, g.COPY_OP_DTYPE)
    want = d0.copy()
    doff, soff = oracle.level_offsets(dbl.boxes, 1, 3), oracle.level_offsets(dbl.boxes, 1, 5)
    for n, o in enumerate(ops):
        b = n // 2                               # two ops per destination box, on different components
        sb = int(rng.integers(0, dbl.size()))
        box = dbl.boxes[b]
        sbox = dbl.boxes[sb]
        ext = rng.integers(1, 5, 3)
        o["dst_box"], o["src_box"] = b, sb
        o["dst_lo"] = box[:3] - 1 + np.array([rng.integers(0, 7 - ext[d]) for d in range(3)])
        o["dst_hi"] = o["dst_lo"] + ext - 1
        o["src_lo"] = sbox[:3] - 1 + np.array([rng.integers(0, 7 - ext[d]) for d in range(3)])
        o["dst_comp"], o["src_comp"], o["ncomp"] = n % 2, int(rng.integers(0, 4)), 1
        o["comp_flags"] = 0xFFFFFFFF if n % 3 else 0x1
        dfl, dfh = box[:3] - 1, box[3:] + 1
        sfl, sfh = sbox[:3] - 1, sbox[3:] + 1
        L.sgco_fab_copy(po.i3(dfl), po.i3(dfh), 3, want[doff[b]:], po.i3(sfl), po.i3(sfh), 5, s0[soff[sb]:],
                        po.i3(o["dst_lo"]), po.i3(o["dst_hi"]), int(o["dst_comp"]), po.i3(o["src_lo"]),
                        po.i3(o["src_lo"] + ext - 1), int(o["src_comp"]), 1, int(o["comp_flags"]))
    cs = g.CopySet(dst, src, ops)
    n0 = g.launch_count()
    cs.run(dst, src)
    assert g.launch_count() - n0 == 1
    assert np.array_equal(dst.download(), want) and np.array_equal(src.download(), s0)
    other = g.LevelData(dbl, 4, 1)
    with pytest.raises(g.SgcError):
        cs.run(other, src)
    ops[0]["dst_lo"][0] -= 5
    with pytest.raises(g.SgcError):
        g.CopySet(dst, src, ops)


def _lb_setup(g, oracle, dlo, dhi, mb, seed):
    rng = np.random.default_rng(seed)
    dbl = g.DisjointBoxLayout(g.Box(dlo, dhi), mb)
    lb = g.LBLevel(dbl)
    fi, mac = oracle.lb_initial(dbl.boxes)
    fi = fi * (1 + 0.05 * rng.standard_normal(fi.size))
    prev = fi * (1 + 0.05 * rng.standard_normal(fi.size))
    oracle.lb_macroscopic(dbl.boxes, fi, mac)
    lb.curr.upload(fi); lb.prev.upload(prev); lb.macro.upload(mac)
    return dbl, lb, fi, prev, mac


def test_lb_phases_vs_oracle(g, oracle):
    """LBPatch::collision, macroscopic, stream and the bounce-back fill of LBLevel::advance, one at a time"""
    for seed, (dlo, dhi, mb) in enumerate([((0, 0, 0), (15, 7, 7), (4, 4, 4)), ((2, -1, 0), (13, 6, 4), (6, 4, 5)),
                                           ((0, 0, 0), (31, 15, 15), (16, 16, 16))]):
        dbl, lb, fi, prev, mac = _lb_setup(g, oracle, dlo, dhi, mb, seed)
        boxes = dbl.boxes
        g.lb_collide(lb.curr, lb.macro)
        oracle.lb_collision(boxes, fi, mac)
        assert np.array_equal(lb.curr.download(), fi)
        lb.curr.exchange(lb.copier)
        items = oracle.copier_items(dlo, dhi, mb, 1, g.LB_PERIODIC, g.LB_TRIM)
        oracle.level_exchange(boxes, 1, 19, items, fi)
        assert np.array_equal(lb.curr.download(), fi)
        lb.bounce.run(lb.curr, lb.curr)
        oracle.lb_bounce(dlo, dhi, boxes, fi)
        assert np.array_equal(lb.curr.download(), fi)
        lb.stream.run(lb.prev, lb.curr)
        oracle.lb_stream(boxes, fi, prev)
        assert np.array_equal(lb.prev.download(), prev) and np.array_equal(lb.curr.download(), fi)
        g.lb_macroscopic(lb.macro, lb.prev)
        oracle.lb_macroscopic(boxes, prev, mac)
        assert np.array_equal(lb.macro.download(), mac)


def test_lb_advance_fused_and_unfused_vs_oracle(g, oracle):
    for seed, (dlo, dhi, mb, nstep) in enumerate([((0, 0, 0), (15, 7, 7), (4, 4, 4), 5), ((2, -1, 0), (13, 6, 4), (6, 4, 5), 4),
                                                  ((0, 0, 0), (31, 15, 15), (16, 16, 16), 3), ((0, 0, 0), (7, 7, 7), (8, 8, 8), 1)]):
        results = []
        for mode in ("phases", "unfused", "fused"):
            dbl, lb, fi, prev, mac = _lb_setup(g, oracle, dlo, dhi, mb, 100 + seed)
            n0 = g.launch_count()
            if mode == "phases":
                for _ in range(nstep):
                    lb.advance_phases()
            else:
                lb.advance(nstep, fuse=(mode == "fused"))
            launches = g.launch_count() - n0
            assert launches == (5 * nstep if mode != "fused" else 3 * nstep + 1), (mode, launches)
            results.append((lb.curr.download(), lb.prev.download(), lb.macro.download()))
        oracle.lb_advance(dlo, dhi, mb, nstep, fi, prev, mac)
        for mode, (c, p, m) in zip(("phases", "unfused", "fused"), results):
            assert np.array_equal(c, fi), mode          # ghost cells included
            assert np.array_equal(m, mac), mode
            if mode != "fused":
                assert np.array_equal(p, prev), mode    # fused: m_prev's valid cells hold post-collision values


def test_step_loops_replayed_as_graphs(g, oracle, monkeypatch):
    """Long runs of sgc_lb_run (fused) and sgc_wave_run replay a captured unit of steps as a CUDA graph (the
    reference's application configurations are so small that a step is launch latency); the bits and the
    launch accounting equal the launch-by-launch loop (SGC_NO_GRAPH=1), and the LB run equals the oracle."""
    dlo, dhi, mb, nstep = (0, 0, 0), (31, 15, 15), (16, 16, 16), 27
    outs = []
    for nograph in (False, True):
        if nograph:
            monkeypatch.setenv("SGC_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("SGC_NO_GRAPH", raising=False)
        dbl, lb, fi, prev, mac = _lb_setup(g, oracle, dlo, dhi, mb, 321)
        n0 = g.launch_count()
        lb.advance(nstep, fuse=True)
        assert g.launch_count() - n0 == 3 * nstep + 1
        outs.append((lb.curr.download(), lb.macro.download()))
    monkeypatch.delenv("SGC_NO_GRAPH", raising=False)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    oracle.lb_advance(dlo, dhi, mb, nstep, fi, prev, mac)
    assert np.array_equal(outs[0][0], fi) and np.array_equal(outs[0][1], mac)
    # wave: three levels rotate, so the unit is three steps; periodic exchange + zero-gradient BC plans
    dlo, dhi, mb = (0, 0, 0), (31, 31, 15), (16, 16, 16)
    dbl, un = make_level(g, dlo, dhi, mb, 1, 1)
    lv = [un, g.LevelData(dbl, 1, 1), g.LevelData(dbl, 1, 1)]
    rng = np.random.default_rng(31)
    init = [rng.standard_normal(un.elems) for _ in range(3)]
    cp = g.Copier().defineExchangeLD(un, 3, 0)
    bc = g.Copier().defineBCZeroGradient(un, g.Box(dlo, dhi))
    res = []
    for nograph in (False, True):
        if nograph:
            monkeypatch.setenv("SGC_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("SGC_NO_GRAPH", raising=False)
        for l, x in zip(lv, init):
            l.upload(x)
        n0 = g.launch_count()
        a, b, c = g.wave_run(lv, cp, bc, 0.0625, 25)
        assert g.launch_count() - n0 == 3 * 25
        res.append((a.download(), b.download(), c.download(), [lv.index(a), lv.index(b), lv.index(c)]))
    monkeypatch.delenv("SGC_NO_GRAPH", raising=False)
    assert res[0][3] == res[1][3]
    for k in range(3):
        assert np.array_equal(res[0][k], res[1][k]), k


def test_lb_matches_reference_golden(g, golden, oracle):
    meta, arr = golden
    for c in meta["lb"]:
        k = c["key"]
        dbl = g.DisjointBoxLayout(g.Box(c["dlo"], c["dhi"]), c["maxbox"])
        for fuse in (False, True):
            lb = g.LBLevel(dbl)
            lb.curr.upload(arr[k + "_fi0"]); lb.macro.upload(arr[k + "_macro0"])
            lb.advance(c["nstep"], fuse=fuse)
            assert np.array_equal(lb.curr.download(), arr[k + "_fi_nofma"]), k    # reference, -ffp-contract=off
            mac = lb.macro.download()
            assert np.array_equal(mac, arr[k + "_macro_nofma"]), k
            ref = arr[k + "_macro"]                                               # reference, default build
            assert np.abs(mac - ref).max() <= 1e-12 * np.abs(ref).max(), k
    for c in meta["lb_hash"]:                    # the shipped configuration from LBLevel::initialData
        dbl = g.DisjointBoxLayout(g.Box(c["dlo"], c["dhi"]), c["maxbox"])
        lb = g.LBLevel(dbl)
        lb.advance(c["nstep"])
        assert "%016x" % oracle.fnv1a64(lb.curr.download()) == c["fnv_fi_nofma"], c["key"]
        assert "%016x" % oracle.fnv1a64(oracle.level_valid_out(dbl.boxes, 1, 4, lb.macro.download())) == c["fnv_macro_nofma"]


def test_plot_gather_matches_reference_writer(g, golden, oracle):
    """LBLevel::writePlotFile -> LevelData::writeCGNSSolData (LevelData.cpp:150-193): the solution values the
    reference hands to CGNS (recorded through the CGNS stand-in) == the device gather of the valid cells."""
    meta, arr = golden
    c = meta["plot"][0]
    dbl = g.DisjointBoxLayout(g.Box(c["dlo"], c["dhi"]), c["maxbox"])
    lb = g.LBLevel(dbl)
    lb.advance(c["nstep"])
    nsol = lb.macro.validElems()
    assert nsol == 4 * 8 * 4 * 8
    want = arr["plot_data"][-nsol:]
    assert np.array_equal(lb.macro.downloadValid(), want)
    # asynchronous form on a transfer stream, ordered after the compute stream by an event
    pin = g.PinnedBuffer(nsol)
    pin.array[:] = -1
    xs, ev, gathered = g.Stream(), g.Event(), g.Event()
    lb.advance(2)
    ev.record()
    xs.wait(ev)
    lb.macro.downloadValid(out=pin, stream=xs, gathered=gathered)
    g.compute_wait(gathered)                     # macro may be rewritten once it has been gathered ...
    lb.advance(1)                                # ... while the D2H copy is still in flight
    xs.sync()
    ref = g.LBLevel(dbl)
    ref.advance(c["nstep"] + 2)
    assert np.array_equal(pin.array, ref.macro.downloadValid())
    # multi-component level with 2 ghost layers and an offset domain
    rng = np.random.default_rng(9)
    dbl2, lev = make_level(g, (1, -1, 1), (12, 10, 12), (4, 4, 4), 3, 2)
    u = rng.standard_normal(lev.elems)
    lev.upload(u)
    assert np.array_equal(lev.downloadValid(), oracle.level_valid_out(dbl2.boxes, 2, 3, u))
