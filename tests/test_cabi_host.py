"""CPU-only checks of the product library: it loads, exports every symbol include/sgc_b200.h
declares, its host-side layout/plan builders agree with the golden vectors from the real reference,
and the data path fails loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def items_as_matrix(it):
    """sgc_motion_item array -> the 24-int record layout of the reference shim / golden vectors
    (local indices and flags are derived by the caller)."""
    m = np.zeros((len(it), 21), np.int32)
    m[:, 0], m[:, 1] = it["recv_box"], it["send_box"]
    m[:, 2:5], m[:, 5:8] = it["recv_lo"], it["recv_hi"]
    m[:, 8:11], m[:, 11:14] = it["src_lo"], it["src_hi"]
    m[:, 14:17] = it["dir"]
    m[:, 17], m[:, 18] = it["recv_proc"], it["send_proc"]
    m[:, 19], m[:, 20] = it["tag_send"], it["tag_recv"]
    return m


def test_library_exports_every_declared_symbol(sgc):
    L = sgc.lib()
    header = open(os.path.join(ROOT, "include", "sgc_b200.h")).read()
    declared = set(re.findall(r"\b(sgc_[a-z0-9_]+)\s*\(", header))
    declared -= {"sgc_status"}
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(L, name), "libsgc_b200.so does not export %s" % name
    assert set(sgc.SYMBOLS) == declared
    assert L.sgc_abi_version() == 1


def test_no_cpu_fallback(sgc):
    import ctypes as C
    L = sgc.lib()
    try:
        sgc.init(0)
        have_gpu = True
    except sgc.SgcError as e:
        have_gpu = False
        assert "no CPU fallback" in str(e)
    if have_gpu:
        pytest.skip("a GPU is present")
    boxes = np.array([[0, 0, 0, 3, 3, 3]], np.int32)
    h = C.c_void_p()
    st = L.sgc_level_create(boxes.ctypes.data, 1, 1, 1, 8, 0, C.byref(h))
    assert st == -2 and b"no CPU fallback" in L.sgc_last_error()
    assert L.sgc_heat_sweep(None, None, 0.1, 1) == -2
    assert L.sgc_exchange(None, None) == -2


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "structuredgridcalc-boxframework_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".H")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in src and "sgcoracle" not in src and "sgcref" not in src, f
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            src = open(os.path.join(dirpath, f), errors="ignore").read()
            assert "sgcoracle" not in src and "sgcref" not in src, f


def test_layout_and_plan_match_reference_golden(sgc, golden):
    meta, arr = golden
    for c in meta["plan"]:
        dbl = sgc.DisjointBoxLayout(sgc.Box(c["dlo"], c["dhi"]), c["maxbox"], c["nproc"], c["rank"])
        name = c["key"][len("plan_"):].rsplit("_n", 1)[0]
        assert np.array_equal(dbl.boxes, arr["layout_" + name])
        it = dbl.copier_items(c["ng"], c["periodic"], c["trim"])
        ref = arr[c["key"]]
        assert len(it) == c["nitem"]
        assert np.array_equal(items_as_matrix(it), ref[:, :21]), c["key"]
        assert np.array_equal(it["recv_box"] - dbl.localIdxBegin(), ref[:, 21])
        assert (it["comp_flags"] == 0xFFFFFFFF).all()


def test_plan_matches_oracle_including_send_regions(sgc, oracle):
    rng = np.random.default_rng(7)
    for _ in range(25):
        mb = rng.integers(2, 6, 3)
        nb = rng.integers(1, 5, 3)
        dlo = rng.integers(-4, 4, 3)
        dhi = dlo + mb * nb - 1
        ng = 1 if mb.min() < 2 else int(rng.integers(1, 3))
        per, trim = int(rng.integers(0, 8)), int(rng.integers(0, 8)) * 2
        nbox = int(np.prod(nb))
        nproc = int(rng.choice([p for p in (1, 2, 4) if nbox % p == 0]))
        for r in range(nproc):
            dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb, nproc, r)
            it = dbl.copier_items(ng, per, trim)
            oi = oracle.copier_items(dlo, dhi, mb, ng, per, trim, nproc, r)
            assert np.array_equal(items_as_matrix(it), oi[:, :21])
            assert np.array_equal(it["send_lo"], oi[:, 24:27]) and np.array_equal(it["send_hi"], oi[:, 27:30])


def test_layout_rejects_bad_decompositions(sgc):
    with pytest.raises(sgc.SgcError):
        sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (62, 63, 63)), 16)          # DisjointBoxLayout.cpp:104-106
    with pytest.raises(sgc.SgcError):
        sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (63, 63, 63)), 16, nproc=7)  # :125


def test_halo_lists_pair_up_across_ranks(sgc):
    """What rank r sends to q must be, item for item, what q expects to receive from r."""
    dom = sgc.Box((0, 0, 0), (15, 15, 31))
    for nproc in (2, 4):
        for per in (0, 7, 4):
            lists = [sgc.DisjointBoxLayout(dom, 4, nproc, r).halo_lists(1, per, 0) for r in range(nproc)]
            for r in range(nproc):
                for peer, recv, send in lists[r]:
                    back = [x for x in lists[peer] if x[0] == r]
                    assert len(back) == 1
                    _, precv, psend = back[0]
                    assert np.array_equal(send, precv)       # same items, same order
                    assert np.array_equal(recv, psend)
                    assert (recv["send_proc"] == peer).all() and (recv["recv_proc"] == r).all()


def test_step_loop_classification(sgc):
    """Which device-resident step loops a plan admits (sgc_plan.cpp: classify_step_loops; host only).  The
    two-step sweep reads its halo from the neighbours' VALID cells: it needs the in-plane neighbours on the same
    rank; z neighbours may be remote (slab partitions); faces without any motion item (physical boundary or
    trimmed) are handled by the boundary instance; edge / corner items are not needed at all."""
    F, LOCAL, SLAB, BND = sgc.LOOP_FUSED, sgc.LOOP_TWO_STEP_LOCAL, sgc.LOOP_TWO_STEP_SLAB, sgc.LOOP_TWO_STEP_BND
    dlo, dhi = (0, 0, 0), (127, 127, 255)
    PX, PY, PZ = 1, 2, 4                       # LayoutIterator.H:35-41
    TRIM_FACE, TRIM_EDGE, TRIM_CORNER = 2, 4, 8
    # fully periodic, one rank: everything local
    assert sgc.step_loop_class(dlo, dhi, 64, 1, 7, 0) == F | LOCAL | SLAB
    # one box, periodic onto itself
    assert sgc.step_loop_class((0, 0, 0), (63, 63, 63), 64, 1, 7, 0) == F | LOCAL | SLAB
    # physical boundaries in some / all directions: the boundary instance, never the plain ones
    for per in (0, PX, PY, PZ, PX | PY, PX | PZ, PY | PZ):
        assert sgc.step_loop_class(dlo, dhi, 64, 1, per, 0) == F | BND, per
    # edge / corner items trimmed (the LB application's plan): the diagonal boxes are reached over two faces
    assert sgc.step_loop_class(dlo, dhi, 64, 1, 7, TRIM_CORNER) == F | LOCAL | SLAB
    assert sgc.step_loop_class(dlo, dhi, 64, 1, 7, TRIM_EDGE | TRIM_CORNER) == F | LOCAL | SLAB
    # z-slab partitions (the reference's ownership rule, DisjointBoxLayout.cpp:121-139): z neighbours remote
    for nproc in (2, 4):
        for rank in range(nproc):
            assert sgc.step_loop_class(dlo, dhi, 64, 1, 7, 0, nproc, rank) == F | SLAB, (nproc, rank)
            got = sgc.step_loop_class(dlo, dhi, 64, 1, PX | PY, 0, nproc, rank)     # physical z ends of the domain:
            want = F | BND if rank in (0, nproc - 1) else F | SLAB                  # only the end ranks see them
            assert got == want, (nproc, rank, got)
    # a partition that cuts a box layer in y (8 ranks x 2 boxes of the 2 x 2 x 4 layout): in-plane neighbours
    # on another rank -> only the one-step loop
    assert sgc.step_loop_class(dlo, dhi, 64, 1, 7, 0, 8, 3) == F


def test_step_loop_pillars(sgc):
    """The pillars the two-step sweep marches down (sgc_plan.cpp: pillar_structure; host only): boxes stacked in z
    that are each other's local z neighbour under the exchange, in the layout's x-fastest order
    (DisjointBoxLayout.cpp:104-139).  Halo planes are then read once per pillar instead of once per box."""
    # BASELINE config 2: 8 x 8 pillars of 8 boxes, whatever the periodicity (a physical boundary is a pillar END)
    for per in (7, 0, 3, 4):
        assert sgc.step_loop_pillars((0, 0, 0), (511, 511, 511), 64, 1, per) == (64, 8), per
    # z-slab partitions keep whole box layers: rank-local pillars
    assert sgc.step_loop_pillars((0, 0, 0), (511, 511, 2047), 64, 1, 7, 0, 4, 1) == (64, 8)
    assert sgc.step_loop_pillars((0, 0, 0), (1023, 1023, 1023), 64, 1, 7, 0, 8, 3) == (256, 2)       # BASELINE config 4
    # one box; one layer of boxes: every box its own pillar
    assert sgc.step_loop_pillars((0, 0, 0), (63, 63, 63), 64) == (1, 1)
    assert sgc.step_loop_pillars((0, 0, 0), (255, 127, 63), 64) == (8, 1)
    # two layers, periodic in z: the wrap-around neighbour of a pillar's top box is its own bottom box
    assert sgc.step_loop_pillars((0, 0, 0), (127, 127, 127), 64) == (4, 2)
    # a partition that cuts a box layer: no pillars
    assert sgc.step_loop_pillars((0, 0, 0), (127, 127, 255), 64, 1, 7, 0, 8, 3) == (2, 1)
    # the z faces trimmed out of the exchange (TrimFace is all or nothing, so every face goes): the boxes above and
    # below are not exchange neighbours and each box keeps its own never-exchanged ghost planes -> box by box
    npil, nbz = sgc.step_loop_pillars((0, 0, 0), (127, 127, 255), 64, 1, 7, 2)
    assert nbz == 1 and npil == 16
    # flat boxes
    assert sgc.step_loop_pillars((0, 0, 0), (127, 127, 31), (64, 64, 8)) == (4, 4)


def test_committed_bench_line_follows_the_contract():
    """profiles/r02_bench_n1.json is bench.py's line from a B200: the keys the driver and the judge read."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    j = json.load(open(os.path.join(root, "profiles", "r02_bench_n1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches",
              "parity", "unfused", "configs"):
        assert k in j, k
    assert j["metric"] == "Gcell-updates/s" and j["dtype"] == "f64" and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"] and j["config"]["workload"].startswith("C2")
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "dram_frac_of_peak", "dram_frac_of_compulsory"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # the dominant kernel advances two steps per launch: its algorithmic bytes are 2 x 16 B per cell, its real DRAM
    # traffic (ncu) is below them, and in real bytes it sits below the copy peak
    assert r["cell_updates_per_cell_per_launch"] == 2 and r["traffic"] < r["algorithmic_bytes_per_launch"]
    assert r["dram_frac_of_compulsory"] < r["dram_frac_of_peak"] < 1.0 < r["frac"]
    c = j["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] == "reference"
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < j["value"]
    assert e["value"] <= e["ceiling_value"] * 1.02 and e["host_bw_ceiling_gbs"] > 0
    assert j["gpu_launches"] > 0 and j["clocks"]["sm_mhz"] > 0
    # parity inside the run: the full-size level against the reference's own CPU run, ghost cells included
    p = j["parity"]
    assert p["ok"] is True and p["against"] == "reference" and p["fnv"] == p["fnv_reference"] and p["ghost_cells_included"]
    assert p["cells_compared"] == 512 * 66 ** 3
    # the other single-GPU configs of BASELINE.json, measured and checked in the same run
    c3, c1 = j["configs"]["C3"], j["configs"]["C1"]
    assert c3["parity"]["ok"] is True and c3["value"] >= 350.0 and c3["motion_items"] == 851968
    assert c3["rebox_factors"] == [4, 4, 4]
    assert j["pillars"] == [64, 8]            # the two-step sweep marched down 8 x 8 pillars of 8 boxes
    assert c1["parity"]["ok"] is True and c1["parity"]["linear_field_B_is_zero"] is True and c1["gpu_launches_per_step"] == 16
    u = j["unfused"]
    assert u["exchange_x_faces_from_edge_strips"] is True and u["exchange_ms_per_sweep"] < 0.1


def test_bench_config_selection(monkeypatch):
    """bench.py --config maps BASELINE.json's configs onto level shapes; C4 divides its 16 box layers among the ranks."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for argv, world, want in ((["bench.py"], 1, ("C2", 512, 64, 0)), (["bench.py"], 8, ("C5", 512, 64, 0)),
                              (["bench.py", "--config", "C3"], 1, ("C3", 512, 16, 0)),
                              (["bench.py", "--config", "C4", "--gpus", "8"], 8, ("C4", 1024, 64, 128))):
        monkeypatch.setattr(sys, "argv", argv)
        monkeypatch.setenv("WORLD_SIZE", str(world))
        a = bench.parse()
        assert (a.config, a.domain, a.box, a.domain_z) == want
        cfg = bench.workload_config(a, world)
        assert cfg["workload"].startswith(a.config) and "model" not in cfg
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "C4"])
    monkeypatch.setenv("WORLD_SIZE", "3")
    with pytest.raises(SystemExit):
        bench.parse()
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "C1"])
    monkeypatch.setenv("WORLD_SIZE", "1")
    assert bench.workload_config(bench.parse(), 1)["workload"].startswith("C1")
    assert bench.global_domain(argparse_ns(domain=512, domain_z=0), 4) == (512, 512, 2048)


def argparse_ns(**kw):
    import argparse
    return argparse.Namespace(**kw)
