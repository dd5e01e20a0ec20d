"""Two-steps-per-pass sweep (sgc_stream2.cu): parity against the oracle, then timing on the C2 level.
    python profiles/micro/t2_probe.py [parity|perf|all] [periodic flags for perf, default 7] [box, default 64]
Test infrastructure: imports oracle/ as the checker only.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from _loadpkg import load_package  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
sgc = load_package()
sgc.init(0)


def run(dlo, dhi, mb, per, f, nstep, contract, u0):
    dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb)
    a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
    a.upload(u0)
    b.upload(u0)
    cp = sgc.Copier().defineExchangeLD(a, per, 0)
    n0 = sgc.launch_count()
    res = sgc.heat_run(a, b, cp, f, nstep, sgc.SGC_FMA if contract else 0)
    out = res.download()
    other = (b if res is a else a).download()
    return out, other, sgc.launch_count() - n0


if what in ("parity", "all"):
    import pyoracle
    O = pyoracle.Oracle()
    ok = True
    cases = [((0, 0, 0), (63, 63, 63), (64, 64, 64), 7, 0.125, 10, 1),
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 7, 0.31, 7, 1),
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 7, 1.0 / 6.0, 12, 0),
             ((0, 0, 0), (191, 63, 127), (64, 64, 64), 7, 0.2, 9, 1),
             ((0, 0, 0), (63, 127, 15), (64, 64, 8), 7, 0.2, 11, 1),
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 3, 0.125, 8, 1),     # physical boundaries
             ((0, 0, 0), (127, 127, 127), (64, 64, 64), 0, 0.2, 9, 1),
             ((0, 0, 0), (191, 127, 63), (64, 64, 64), 4, 1.0 / 6.0, 8, 0),
             ((0, 0, 0), (63, 191, 127), (64, 64, 64), 1, 0.3, 10, 1),
             ((0, 0, 0), (127, 127, 15), (64, 64, 8), 2, 0.3, 7, 1),
             # 32^2 boxes (8 warps, three CTAs per SM)
             ((0, 0, 0), (31, 31, 31), (32, 32, 32), 7, 0.2, 10, 1),
             ((0, 0, 0), (95, 63, 63), (32, 32, 32), 7, 1.0 / 6.0, 9, 0),
             ((0, 0, 0), (63, 95, 15), (32, 32, 8), 7, 0.3, 12, 1),
             ((0, 0, 0), (63, 63, 63), (32, 32, 32), 0, 0.2, 8, 1),
             ((0, 0, 0), (95, 63, 31), (32, 32, 16), 5, 0.3, 7, 1),
             # 16^2 boxes (4 warps, two roles per warp; experimental: SGC_T2_16=1)
             ((0, 0, 0), (15, 15, 15), (16, 16, 16), 7, 0.2, 10, 1),
             ((0, 0, 0), (63, 47, 31), (16, 16, 16), 7, 1.0 / 6.0, 9, 0),
             ((0, 0, 0), (31, 47, 15), (16, 16, 8), 0, 0.3, 12, 1),
             ((0, 0, 0), (47, 31, 31), (16, 16, 16), 6, 0.3, 7, 1)]
    for (dlo, dhi, mb, per, f, nstep, contract) in cases:
        dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb)
        rng = np.random.default_rng(3)
        probe = sgc.LevelData(dbl, 1, 1)
        u0 = rng.standard_normal(probe.elems)
        del probe
        t = time.time()
        out, other, nl = run(dlo, dhi, mb, per, f, nstep, contract, u0)
        want = u0.copy()
        O.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want, contract)
        good = np.array_equal(out, want)
        good_prev = True      # (the other level's ghost cells are pinned by tests/test_gpu_parity.py)
        nbad = int((out != want).sum())
        print("parity", dhi, mb, "per", per, "f", f, "nstep", nstep, "fma", contract, "launches", nl,
              "OK" if good else "MISMATCH (%d of %d)" % (nbad, out.size),
              "%.1fs" % (time.time() - t), flush=True)
        if not good:
            bad = np.nonzero(out != want)[0]
            print("   first bad flat indices", bad[:8], "got", out[bad[:3]], "want", want[bad[:3]])
        ok = ok and good and good_prev
    # known answer: 64^3 periodic, hash init, 100 sweeps, f = 0.125
    dlo, dhi, mb = (0, 0, 0), (63, 63, 63), (64, 64, 64)
    dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb)
    u0 = O.fill_hash(dbl.boxes, 1, 0.0)
    out, _, nl = run(dlo, dhi, mb, 7, 0.125, 100, 1, u0)
    h = "%016x" % O.fnv1a64(O.gather(dlo, dhi, dbl.boxes, 1, 1, 0, out))
    print("known answer 64^3 x 100:", h, "launches", nl, "OK" if h == "978934eb8172d8e9" else "MISMATCH", flush=True)
    ok = ok and h == "978934eb8172d8e9"
    print("PARITY", "GREEN" if ok else "RED", flush=True)

if what in ("perf", "all"):
    n, sweeps = 512, 100
    box = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
    a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
    rng = np.random.default_rng(0)
    a.upload(rng.random(a.elems))
    b.copyFrom(a)
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 7      # periodic flags (7 = xyz; 0 = physical boundaries all round)
    cp = sgc.Copier().defineExchangeLD(a, per, 0)
    for _ in range(2):
        sgc.heat_run(a, b, cp, 0.125, sweeps)
    sgc.sync()
    sgc.profile_enable(True)
    for k in range(3):
        sgc.profile_read(k)
    e0, e1 = sgc.Event(), sgc.Event()
    K = 3
    e0.record()
    for _ in range(K):
        sgc.heat_run(a, b, cp, 0.125, sweeps)
    e1.record()
    sgc.sync()
    ms = e0.elapsed_ms(e1) / K
    n0, t0 = sgc.profile_read(0)
    n2, t2 = sgc.profile_read(2)
    cells = float(n) ** 3
    print(json.dumps({"box": box, "periodic": per, "ms_per_100_sweeps": ms, "gcell_per_s": cells * sweeps / (ms * 1e-3) / 1e9,
                      "one_step_launches": n0, "one_step_ms": t0 / max(n0, 1),
                      "two_step_launches": n2, "two_step_ms": t2 / max(n2, 1),
                      "two_step_gcell_per_s": 2 * cells / (t2 / max(n2, 1) * 1e-3) / 1e9 if n2 else None}), flush=True)
