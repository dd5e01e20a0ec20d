"""A handful of small two-step-sweep cases (for a plain run, or under compute-sanitizer where the pool allows it): pillars of two boxes with a periodic wrap,
physical boundaries, flat boxes, the 32^2 instance, a single box, and a re-boxed level of 16^3 boxes.
    compute-sanitizer --tool memcheck python profiles/micro/t2_memcheck.py
Test infrastructure: imports oracle/ as the checker only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from _loadpkg import load_package  # noqa: E402
import pyoracle  # noqa: E402

sgc = load_package()
sgc.init(0)
O = pyoracle.Oracle()
cases = [((0, 0, 0), (127, 127, 127), (64, 64, 64), 7, 0.2, 8),
         ((0, 0, 0), (63, 127, 191), (64, 64, 64), 7, 0.125, 7),      # pillars of three boxes
         ((0, 0, 0), (127, 127, 127), (64, 64, 64), 3, 0.2, 8),       # physical z ends (boundary instance)
         ((0, 0, 0), (127, 63, 63), (64, 64, 64), 0, 0.2, 6),
         ((0, 0, 0), (63, 127, 31), (64, 64, 8), 7, 0.3, 9),          # flat boxes: pillars of four
         ((0, 0, 0), (63, 63, 63), (32, 32, 32), 7, 0.2, 8),
         ((0, 0, 0), (63, 63, 63), (64, 64, 64), 7, 0.125, 6),
         ((0, 0, 0), (63, 63, 127), (16, 16, 16), 7, 0.125, 12)]      # re-boxed: unions of 4 x 4 x 4 boxes
ok = True
for dlo, dhi, mb, per, f, nstep in cases:
    dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb)
    a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
    u0 = np.random.default_rng(5).standard_normal(a.elems)
    a.upload(u0)
    b.upload(u0)
    cp = sgc.Copier().defineExchangeLD(a, per, 0)
    res = sgc.heat_run(a, b, cp, f, nstep)
    got = res.download()
    want = u0.copy()
    O.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want, 1)
    good = bool(np.array_equal(got, want))
    print(dhi, mb, "per", per, "nstep", nstep, "pillars", cp.pillars(), "OK" if good else "MISMATCH", flush=True)
    ok = ok and good
sgc.sync()
print("MEMCHECK CASES", "GREEN" if ok else "RED", flush=True)
