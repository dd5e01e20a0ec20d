import os, sys
import numpy as np
ROOT = os.getcwd()
sys.path.insert(0, ROOT)
from _loadpkg import load_package
sgc = load_package(); sgc.init(0)
n, box = 512, 16
dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
a.upload(np.random.default_rng(0).random(a.elems)); b.copyFrom(a)
cp = sgc.Copier().defineExchangeLD(a, 7, 0)
sgc.heat_run(a, b, cp, 0.125, 12); sgc.sync()
