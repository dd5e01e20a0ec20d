"""Launch-bound step loops on the reference's own (small) application configurations, with and without the
CUDA-graph replay (SGC_NO_GRAPH=1):  latticeBoltzmann 64 x 32 x 32 in 16^3 boxes; wave on a single 64^3 box.
    python profiles/micro/graph_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402
sgc = load_package(); sgc.init(0)
for nograph in (False, True):
    if nograph:
        os.environ["SGC_NO_GRAPH"] = "1"
    else:
        os.environ.pop("SGC_NO_GRAPH", None)
    dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (63, 31, 31)), (16, 16, 16))
    lb = sgc.LBLevel(dbl); lb.advance(30, fuse=True); sgc.sync()
    e0, e1 = sgc.Event(), sgc.Event(); e0.record(); lb.advance(2000, fuse=True); e1.record(); sgc.sync()
    ms = e0.elapsed_ms(e1)
    print("graph %-3s LB 64x32x32/16^3: %.2f us per step, %.2f Gcell/s" % ("off" if nograph else "on", ms / 2000 * 1e3, 65536 * 2000 / (ms * 1e-3) / 1e9), flush=True)
    h = 64
    box = [[0, 0, 0, h - 1, h - 1, h - 1]]
    lv = [sgc.LevelData(None, 1, 1, boxes=box) for _ in range(3)]
    rng = np.random.default_rng(0)
    for l in lv:
        l.upload(rng.random(l.elems))
    bc = sgc.Copier().defineBCZeroGradient(lv[0], sgc.Box((0, 0, 0), (h - 1,) * 3))
    sgc.wave_run(lv, None, bc, 0.0625, 30); sgc.sync()
    e0, e1 = sgc.Event(), sgc.Event(); e0.record(); sgc.wave_run(lv, None, bc, 0.0625, 3000); e1.record(); sgc.sync()
    ms = e0.elapsed_ms(e1)
    print("graph %-3s wave 64^3 single box: %.2f us per step, %.2f Gcell/s" % ("off" if nograph else "on", ms / 3000 * 1e3, h ** 3 * 3000 / (ms * 1e-3) / 1e9), flush=True)
