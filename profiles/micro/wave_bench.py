"""Times the leap-frog wave step (sgc_wave_step, 24 algorithmic bytes per cell-update)
on the C2 level shape: 512^3 in 64^3 boxes, periodic.  One step = exchange(un) + wave_step + rotate.
    python profiles/micro/wave_bench.py [box]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402

sgc = load_package()
sgc.init(0)
box = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 512
dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
lv = [sgc.LevelData(dbl, 1, 1) for _ in range(3)]
rng = np.random.default_rng(0)
u = rng.random(lv[0].elems)
lv[0].upload(u)
lv[2].copyFrom(lv[0])
lv[1].setVal(0.0)
cp = sgc.Copier().defineExchangeLD(lv[0], 7, 0)
cur, nxt, old = 0, 1, 2
for _ in range(5):
    lv[cur].exchange(cp)
    sgc.wave_step(lv[nxt], lv[cur], lv[old], 0.0625)
    cur, nxt, old = nxt, old, cur
sgc.sync()
sgc.profile_enable(True)
sgc.profile_read(0), sgc.profile_read(1)
e0, e1 = sgc.Event(), sgc.Event()
K = 50
e0.record()
for _ in range(K):
    lv[cur].exchange(cp)
    sgc.wave_step(lv[nxt], lv[cur], lv[old], 0.0625)
    cur, nxt, old = nxt, old, cur
e1.record()
sgc.sync()
ms = e0.elapsed_ms(e1) / K
nsw, sw = sgc.profile_read(0)
nex, ex = sgc.profile_read(1)
cells = float(n) ** 3
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"op": "wave_step", "box": box, "ms_per_step": ms, "gcell_per_s": cells / (ms * 1e-3) / 1e9,
                  "wave_kernel_ms": sw / nsw, "exchange_ms": ex / nex,
                  "wave_kernel_gbs_24B": cells * 24 / (sw / nsw * 1e-3) / 1e9,
                  "frac_of_measured_hbm": cells * 24 / (sw / nsw * 1e-3) / 1e9 / peak}))
