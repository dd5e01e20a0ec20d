"""The C++ facade (include/BoxFramework/*.H, same class names as the reference) above the C ABI.

* The REFERENCE's own serial test programs — compiled UNMODIFIED from /root/reference against this
  repo's headers by apps/build_apps.sh — must print "passed" with every BaseFab / LevelData /
  exchange operation executed on the GPU.
* The repo's drivers written against the facade (apps/laplacian.cpp, apps/levelheat.cpp) must
  reproduce the reference's known answers.
The binaries are built in the build container (they travel with the snapshot)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "apps", "_build")

REF_TESTS = ["testIntVect", "testBox", "testBoxIterator", "testBaseFab", "testDisjointBoxLayout",
             "testLayoutIterator", "testLevelData"]
HOST_ONLY = {"testIntVect", "testBox", "testDisjointBoxLayout", "testLayoutIterator"}


def _run(exe, *args, timeout=300):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        pytest.skip("%s not built (apps/build_apps.sh needs /root/reference for ref_* programs)" % exe)
    r = subprocess.run([path] + list(args), capture_output=True, text=True, timeout=timeout, cwd=BUILD)
    return r


@pytest.mark.parametrize("name", sorted(HOST_ONLY))
def test_reference_host_tests_against_facade(name):
    r = _run("ref_" + name)
    assert r.returncode == 0 and "passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", [t for t in REF_TESTS if t not in HOST_ONLY])
def test_reference_device_tests_against_facade(name):
    r = _run("ref_" + name)
    assert r.returncode == 0 and "passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_vlaloops_address_identities_against_facade():
    """application/vlaloops/vlaloops.cpp, compiled unmodified: &vla[MD_IX(i, 0)] == &fab(IntVect(i0,i1,i2), 0) for
    the plain VLA view, the offset-subtracted view and the MD_ARRAY / MD_BOXLOOP macros (asserts live)."""
    r = _run("ref_vlaloops")
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_laplacian_driver_facade():
    for n in ("64", "128"):                       # as shipped / BASELINE config 1
        r = _run("laplacian", n, "16")
        assert r.returncode == 0, r.stdout + r.stderr
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert line["nonzero_cells"] == 0 and line["n"] == int(n)


@pytest.mark.gpu
def test_levelheat_driver_facade_known_answer(oracle):
    # 64^3 periodic, hash init, f = 0.125, 100 steps: the local hash of a 1-rank run covers all boxes
    import numpy as np
    for bs in ("16", "32"):
        r = _run("levelheat", "64", bs, "100")
        assert r.returncode == 0, r.stdout + r.stderr
        line = json.loads(r.stdout.strip().splitlines()[-1])
        dlo, dhi, mb = (0, 0, 0), (63, 63, 63), (int(bs),) * 3
        boxes, _, _ = oracle.layout(dlo, dhi, mb)
        u = oracle.fill_hash(boxes, 1, 0.0)
        oracle.level_heat(dlo, dhi, mb, 1, 7, 0, 0.125, 100, u, 1)
        # same traversal as the driver: box by box, valid cells, x fastest
        off, m, h = 0, int(bs) + 2, 1469598103934665603
        chunks = []
        for _ in range(len(boxes)):
            chunks.append(u[off:off + m ** 3].reshape(m, m, m)[1:-1, 1:-1, 1:-1].copy().ravel())
            off += m ** 3
        want = "%016x" % oracle.fnv1a64(np.concatenate(chunks))
        assert line["local_fnv1a64"] == want


@pytest.mark.gpu
def test_wave_driver_facade_matches_shipped_wavepatch(oracle, golden):
    # apps/wave (C++ facade) against the golden vector produced by the reference's real WavePatch:
    # same initial pulse, 10 leap-frog steps with zero-gradient BC, valid cells bit for bit
    import numpy as np
    meta, arr = golden
    c = [x for x in meta["wavepatch"] if x["key"] == "wp16_pulse"][0]
    r = _run("wave", str(c["h"]), str(c["nstep"]))
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    m = c["h"] + 2
    want = arr["wp16_pulse_un"].reshape(m, m, m)[1:-1, 1:-1, 1:-1].copy()
    assert line["un_valid_fnv1a64"] == "%016x" % oracle.fnv1a64(want)


@pytest.mark.gpu
def test_lattice_boltzmann_driver_facade_known_answer(golden, tmp_path):
    # apps/latticeBoltzmann (C++ facade, LBLevel interface of the reference) on the shipped
    # configuration: hash of the plot data (valid cells of the macroscopic level) after 1 and 40
    # steps == the reference application's (-ffp-contract=off build), in one call and step by step
    import numpy as np
    meta, arr = golden
    for c in meta["lb_hash"]:
        for per_call in ("0", "1", "7"):
            path = os.path.join(BUILD, "latticeBoltzmann")
            if not os.path.exists(path):
                pytest.skip("latticeBoltzmann not built")
            r = subprocess.run([path, "64", "32", "32", "16", str(c["nstep"]), "0", per_call], capture_output=True,
                               text=True, timeout=300, cwd=tmp_path)
            assert r.returncode == 0, r.stdout + r.stderr
            line = json.loads(r.stdout.strip().splitlines()[-1])
            assert line["macro_fnv1a64"] == c["fnv_macro_nofma"], (c["key"], per_call)
            assert abs(line["mass"] - c["mass_nofma"]) <= 1e-9 * c["mass_nofma"]
    # plot file: header + the gathered valid cells
    r = subprocess.run([os.path.join(BUILD, "latticeBoltzmann"), "8", "4", "8", "4", "3", "3"], capture_output=True,
                       text=True, timeout=300, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(os.path.join(tmp_path, "plot", "solution0000.sgcplot"), "rb").read()
    head, _, rest = raw.partition(b"data ")
    n = int(rest.split(b"\n", 1)[0])
    data = np.frombuffer(rest.split(b"\n", 1)[1], np.float64)
    assert n == data.size == 4 * 8 * 4 * 8 and head.startswith(b"SGCPLOT 1\nzones 4\n")
    assert (data[:64] == 1.0).all() and (data[64:256] == 0.0).all()     # initialData: rho = 1, u = 0 (zone 1)
