"""The oracle (oracle/sgc_oracle.c) against the committed golden vectors, which were generated from
the REAL reference library by oracle/gen_golden.py.  CPU only."""
import numpy as np
import pytest


def test_plan_items_match_reference_golden(oracle, golden):
    meta, arr = golden
    assert len(meta["plan"]) >= 20
    for c in meta["plan"]:
        items = oracle.copier_items(c["dlo"], c["dhi"], c["maxbox"], c["ng"], c["periodic"], c["trim"],
                                    c["nproc"], c["rank"])
        ref = arr[c["key"]]
        assert len(items) == c["nitem"] == len(ref), c["key"]
        assert np.array_equal(items[:, :24], ref), c["key"]


def test_plan_counts_survey_A3(oracle):
    # SURVEY A.3 [probe]: item counts for ng=1, bs=4: nb=2: 56/208, nb=3: 316/702, nb=4: 936/1664, nb=8: 10136/13312
    want = {2: (56, 208, 48, 144, 24, 48), 3: (316, 702, 252, 486, 108, 162),
            4: (936, 1664, 720, 1152, 288, 384), 8: (10136, 13312, 7392, 9216, 2688, 3072)}
    for nb, w in want.items():
        hi = (4 * nb - 1,) * 3
        got = []
        for trim in (0, 8, 12):
            for per in (0, 7):
                got.append(len(oracle.copier_items((0, 0, 0), hi, (4, 4, 4), 1, per, trim)))
        assert tuple(got) == w, (nb, got)


def test_layout_matches_reference_golden(oracle, golden):
    meta, arr = golden
    seen = set()
    for c in meta["plan"]:
        name = c["key"][len("plan_"):].rsplit("_n", 1)[0]
        if name in seen:
            continue
        seen.add(name)
        boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"], 1)
        assert np.array_equal(boxes, arr["layout_" + name]), name


def test_layout_ownership_blocks(oracle):
    # DisjointBoxLayout.cpp:121-139: rank k owns [k*bpp,(k+1)*bpp) of the x-fastest box list
    boxes, owner, nb = oracle.layout((0, 0, 0), (63, 63, 127), (16, 16, 16), 8)
    assert nb == (4, 4, 8) and len(boxes) == 128
    assert np.array_equal(owner, np.arange(128) // 16)
    assert tuple(boxes[1]) == (16, 0, 0, 31, 15, 15) and tuple(boxes[4]) == (0, 16, 0, 15, 31, 15)
    with pytest.raises(ValueError):
        oracle.layout((0, 0, 0), (63, 63, 63), (16, 16, 16), 7)      # must divide (:125)
    with pytest.raises(ValueError):
        oracle.layout((0, 0, 0), (62, 63, 63), (16, 16, 16), 1)      # boxes must fit (:104-106)


def test_exchange_matches_reference_golden(oracle, golden):
    meta, arr = golden
    for c in meta["exchange"]:
        boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"])
        items = oracle.copier_items(c["dlo"], c["dhi"], c["maxbox"], c["ng"], c["periodic"], c["trim"])
        assert len(items) == c["nitem"]
        u = arr["xin_" + c["key"]].copy()
        skipped = oracle.level_exchange(boxes, c["ng"], c["ncomp"], items, u)
        assert skipped == 0
        assert np.array_equal(u, arr["xout_" + c["key"]]), c["key"]


def test_heat_matches_reference_golden(oracle, golden):
    meta, arr = golden
    for c in meta["heat"]:
        boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"])
        u = oracle.fill_hash(boxes, 1, 0.0)
        oracle.level_heat(c["dlo"], c["dhi"], c["maxbox"], 1, c["periodic"], 0, c["f"], c["nstep"], u, c["contract"])
        assert np.array_equal(u, arr["heat_" + c["key"]]), c["key"]
        g = oracle.gather(c["dlo"], c["dhi"], boxes, 1, 1, 0, u)
        assert "%016x" % oracle.fnv1a64(g) == c["fnv"]


def test_heat_known_answer_hashes(oracle, golden):
    meta, arr = golden
    by = {c["key"]: c for c in meta["heat_hash"]}
    # SURVEY §8c [probe]: identical for box sizes 16/32/64 and for contract fast|off at f=0.125
    assert by["k64_b16"]["fnv"] == by["k64_b32"]["fnv"] == by["k64_b64"]["fnv"] == "978934eb8172d8e9"
    assert by["k64_b16_sixth_fma"]["fnv"] == "fb0472d7beb7e3a6"
    assert by["k64_b16_sixth_nofma"]["fnv"] == "1d09ec6f6796b0be"
    for key in ("k64_b16", "k64_b64", "k64_b16_sixth_fma", "k64_b16_sixth_nofma"):
        c = by[key]
        boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"])
        u = oracle.fill_hash(boxes, 1, 0.0)
        n, _ = oracle.level_heat(c["dlo"], c["dhi"], c["maxbox"], 1, c["periodic"], 0, c["f"], c["nstep"], u, c["contract"])
        assert n == c["nitem"]
        g = oracle.gather(c["dlo"], c["dhi"], boxes, 1, 1, 0, u)
        assert "%016x" % oracle.fnv1a64(g) == c["fnv"], key
        assert "%016x" % oracle.fnv1a64(u) == c["fnv_level"], key
    # power-of-two coefficient: contraction-invariant
    c = by["k64_b16"]
    boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"])
    u = oracle.fill_hash(boxes, 1, 0.0)
    oracle.level_heat(c["dlo"], c["dhi"], c["maxbox"], 1, 7, 0, 0.125, 100, u, 0)
    assert "%016x" % oracle.fnv1a64(oracle.gather(c["dlo"], c["dhi"], boxes, 1, 1, 0, u)) == c["fnv"]


def test_laplacian_matches_reference_golden(oracle, golden):
    meta, arr = golden
    for c in meta["laplacian"]:
        n = c["n"]
        if c["init"] == "lin":
            ii = np.arange(1, n + 1, dtype=np.float64)
            A = (ii[None, None, :] + ii[None, :, None] + ii[:, None, None]).ravel().copy()
        else:
            A = oracle.fill_hash(np.array([[1, 1, 1, n, n, n]], np.int32), 0)
        B = oracle.laplacian(n, c["niter"], A)
        assert np.array_equal(B, arr[c["key"]]), c["key"]
        if c["init"] == "lin":
            assert not B.any()          # laplacian.cpp:80-86,115-120: Laplacian of i+j+k == 0 exactly


def test_wave_matches_reference_golden(oracle, golden):
    meta, arr = golden
    for c in meta["wave"]:
        boxes = np.array([list(c["dlo"]) + list(c["dhi"])], np.int32)
        u0 = oracle.fill_hash(boxes, 1, 0.0)
        u1 = u0.copy()
        oracle.wave(c["dlo"], c["dhi"], c["factor"], c["nstep"], u0, u1, c["contract"])
        assert np.array_equal(u0, arr[c["key"] + "_u0"]), c["key"]
        assert np.array_equal(u1, arr[c["key"] + "_u1"]), c["key"]


def _valid(a, h):
    m = h + 2
    return a.reshape(m, m, m)[1:-1, 1:-1, 1:-1]


def test_shipped_wavepatch_matches_reference_golden(oracle, golden):
    # WavePatch::advance as shipped (USE_VEX pencil path, MD_DIRSUM form): valid cells, bit for bit
    meta, arr = golden
    assert len(meta["wavepatch"]) >= 3
    for c in meta["wavepatch"]:
        h = c["h"]
        u0, u1 = arr[c["key"] + "_u_init"].copy(), arr[c["key"] + "_um1_init"].copy()
        oracle.wave_dirsum((0, 0, 0), (h - 1,) * 3, c["factor"], c["nstep"], u0, u1, 1)
        assert np.array_equal(_valid(u0, h), _valid(arr[c["key"] + "_un"], h)), c["key"]
        assert np.array_equal(_valid(u1, h), _valid(arr[c["key"] + "_unm1"], h)), c["key"]


# ------------------------------------------------------------------ lattice Boltzmann application
def test_lb_advance_matches_reference_golden(oracle, golden):
    """LBLevel::advance (LBLevel.cpp:7-127): bit-exact (ghost cells included) against the reference built
    with -ffp-contract=off, <= 1e-12 relative against its default (contracting) build."""
    meta, arr = golden
    assert len(meta["lb"]) >= 2
    for c in meta["lb"]:
        k = c["key"]
        boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"])
        fi, mac = arr[k + "_fi0"].copy(), arr[k + "_macro0"].copy()
        prev = oracle.lb_initial(boxes)[0]
        oracle.lb_advance(c["dlo"], c["dhi"], c["maxbox"], c["nstep"], fi, prev, mac)
        assert np.array_equal(fi, arr[k + "_fi_nofma"]), k
        assert np.array_equal(mac, arr[k + "_macro_nofma"]), k
        ref = arr[k + "_macro"]
        assert np.abs(mac - ref).max() <= 1e-12 * np.abs(ref).max(), k


def test_lb_shipped_configuration_hashes(oracle, golden):
    """latticeBoltzmann.cpp:16,36: 64x32x32 in 16^3 boxes from LBLevel::initialData"""
    meta, arr = golden
    for c in meta["lb_hash"]:
        boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"])
        fi, mac = oracle.lb_initial(boxes)
        prev = fi.copy()
        oracle.lb_advance(c["dlo"], c["dhi"], c["maxbox"], c["nstep"], fi, prev, mac)
        assert "%016x" % oracle.fnv1a64(fi) == c["fnv_fi_nofma"], c["key"]
        assert "%016x" % oracle.fnv1a64(oracle.level_valid_out(boxes, 1, 4, mac)) == c["fnv_macro_nofma"], c["key"]


def test_plot_record_matches_reference_golden(oracle, golden):
    """What LBLevel::writePlotFile hands to CGNS (DisjointBoxLayout.cpp:224-383, LevelData.cpp:65-196):
    per zone the vertex coordinates, then per zone and variable the valid cells, ghosts stripped."""
    meta, arr = golden
    c = meta["plot"][0]
    boxes, owner, nb = oracle.layout(c["dlo"], c["dhi"], c["maxbox"])
    fi, mac = oracle.lb_initial(boxes)
    prev = fi.copy()
    oracle.lb_advance(c["dlo"], c["dhi"], c["maxbox"], c["nstep"], fi, prev, mac)
    data, pm, names = arr["plot_data"], arr["plot_meta"], c["names"]
    nbox = len(boxes)
    assert names[:nbox] == ["Zone_%06d" % (b + 1) for b in range(nbox)]
    assert names[nbox:nbox + 3] == ["COMPONENT_X", "COMPONENT_Y", "COMPONENT_Z"]
    assert names[4 * nbox:4 * nbox + 4] == ["density", "x-velocity", "y-velocity", "z-velocity"]
    # coordinates: (n+1)^3 vertices of every box, value = the index along the direction
    pos = 0
    for b in boxes:
        nv = b[3:] - b[:3] + 2
        for d in range(3):
            idx = np.indices((nv[2], nv[1], nv[0]))[2 - d] + b[d]
            assert np.array_equal(data[pos:pos + idx.size], idx.ravel().astype(float))
            pos += idx.size
    # solution: the valid cells of every box and component
    assert np.array_equal(data[pos:], oracle.level_valid_out(boxes, 1, 4, mac))
