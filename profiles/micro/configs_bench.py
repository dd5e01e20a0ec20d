"""One number per single-GPU BASELINE config next to bench.py's C2 line.
  C1: reference application driver, single 128^3 double box, 16 iterations of the 3-direction
      Laplacian accumulate (laplacian.cpp:90-110), un-fused (16 launches) and fused (1 launch);
      24 algorithmic bytes per cell-iteration un-fused.
  C3: 32 768 boxes of 16^3 (512^3 domain), periodic, 100 x {exchange; sweep}; plan build time and
      launches per step loop.
    python profiles/micro/configs_bench.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402

sgc = load_package()
sgc.init(0)
peak = 6460.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

# ---- C1
n = 128
A = sgc.LevelData(None, 1, 0, boxes=[[1, 1, 1, n, n, n]])
B = sgc.LevelData(None, 1, 0, boxes=[[2, 2, 2, n - 1, n - 1, n - 1]])
rng = np.random.default_rng(0)
A.upload(rng.random(A.elems))
cells = float(n - 2) ** 3
for fuse in (0, 1):
    B.setVal(0.0)
    for _ in range(3):
        sgc.laplacian_apply(B, A, 0.5, 16, fuse)
    sgc.sync()
    e0, e1 = sgc.Event(), sgc.Event()
    K = 50
    e0.record()
    for _ in range(K):
        sgc.laplacian_apply(B, A, 0.5, 16, fuse)
    e1.record()
    sgc.sync()
    ms = e0.elapsed_ms(e1) / K
    print(json.dumps({"config": "C1", "box": n, "iterations": 16, "fused_iterations": bool(fuse), "ms_per_16_iterations": ms,
                      "gcell_updates_per_s": cells * 16 / (ms * 1e-3) / 1e9,
                      "algorithmic_GBs_24B": cells * 16 * 24 / (ms * 1e-3) / 1e9,
                      "note": "16 MB of data: L2-resident, so the algorithmic rate may exceed the HBM peak of %.0f GB/s" % peak}), flush=True)
del A, B

# ---- C3
n, box, sweeps = 512, 16, 100
dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
a.upload(rng.random(a.elems))
b.copyFrom(a)
t0 = time.time()
cp = sgc.Copier().defineExchangeLD(a, 7, 0)
plan_ms = (time.time() - t0) * 1e3
for _ in range(2):
    sgc.heat_run(a, b, cp, 0.125, sweeps)
sgc.sync()
l0 = sgc.launch_count()
e0, e1 = sgc.Event(), sgc.Event()
K = 3
e0.record()
for _ in range(K):
    sgc.heat_run(a, b, cp, 0.125, sweeps)
e1.record()
sgc.sync()
ms = e0.elapsed_ms(e1) / K
cells = float(n) ** 3
print(json.dumps({"config": "C3", "boxes": dbl.size(), "box": box, "motion_items": cp.numMotionItem(), "plan_build_ms": plan_ms,
                  "ms_per_100_sweeps": ms, "gcell_updates_per_s": cells * sweeps / (ms * 1e-3) / 1e9,
                  "algorithmic_GBs_16B": cells * sweeps * 16 / (ms * 1e-3) / 1e9,
                  "frac_of_measured_hbm": cells * sweeps * 16 / (ms * 1e-3) / 1e9 / peak,
                  "launches_per_100_sweeps": (sgc.launch_count() - l0) / K}), flush=True)
