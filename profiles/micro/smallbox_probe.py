"""Step loop on narrow boxes (512^3 level in 16^3 / 32^3 boxes, periodic, 100 sweeps).
    python profiles/micro/smallbox_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402
sgc = load_package(); sgc.init(0)
rng = np.random.default_rng(0)
for box in (16, 32):
    n, sweeps = 512, 100
    dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
    a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
    a.upload(rng.random(a.elems)); b.copyFrom(a)
    cp = sgc.Copier().defineExchangeLD(a, 7, 0)
    sgc.heat_run(a, b, cp, 0.125, sweeps); sgc.sync()
    e0, e1 = sgc.Event(), sgc.Event()
    e0.record()
    for _ in range(2):
        sgc.heat_run(a, b, cp, 0.125, sweeps)
    e1.record(); sgc.sync()
    ms = e0.elapsed_ms(e1) / 2
    print("box %d: %.2f ms per 100 sweeps, %.1f Gcell/s" % (box, ms, n ** 3 * sweeps / (ms * 1e-3) / 1e9), flush=True)
    del a, b, cp
