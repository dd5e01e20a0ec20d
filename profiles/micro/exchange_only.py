"""C2 level: sweep, then the stand-alone exchange (strip-sourced x faces) a few times — the target of
`ncu -k regex:k_copy_items`.  SGC_NO_XE=1 gives the plain table."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402

sgc = load_package()
sgc.init(0)
n, box = int(sys.argv[1]) if len(sys.argv) > 1 else 512, int(sys.argv[2]) if len(sys.argv) > 2 else 64
dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
u = np.random.default_rng(1).standard_normal(a.elems)
a.upload(u)
cp = sgc.Copier().defineExchangeLD(a, 7, 0)
for _ in range(3):
    sgc.heat_sweep(b, a, 0.125)
    b.exchange(cp)
    sgc.heat_sweep(a, b, 0.125)
    a.exchange(cp)
sgc.sync()
print("ok", a.edgeStrips(), b.edgeStrips())
