"""Which part of the stand-alone exchange is slow?  The C2 level's motion items split by direction class, each class as
its own item table (sgc_plan_create_items), timed alone: z faces are contiguous 32 KB planes on both sides, y faces are
64 rows of 512 bytes 33 KB apart on both sides, x faces are single elements per row on the read side (no strips here:
these tables are not ghost-only plans) and dense 512-byte strips on the write side.
    python profiles/micro/exchange_directions.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402

sgc = load_package()
sgc.init(0)
n, box = 512, 64
dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
a = sgc.LevelData(dbl, 1, 1)
a.upload(np.random.default_rng(1).standard_normal(a.elems))
items = dbl.copier_items(1, 7, 0)
d = items["dir"]
codim = np.abs(d).sum(axis=1)
classes = {"z faces": (codim == 1) & (d[:, 2] != 0), "y faces": (codim == 1) & (d[:, 1] != 0), "x faces (plain)": (codim == 1) & (d[:, 0] != 0),
           "edges": codim == 2, "corners": codim == 3, "all": codim > 0}
out = {}
for name, mask in classes.items():
    sub = items[mask]
    cp = sgc.Copier().defineItems(a, sub)
    cells = int(np.prod(sub["recv_hi"] - sub["recv_lo"] + 1, axis=1).sum())
    for _ in range(3):
        a.exchange(cp)
    e0, e1 = sgc.Event(), sgc.Event()
    sgc.sync()
    e0.record()
    for _ in range(20):
        a.exchange(cp)
    e1.record()
    sgc.sync()
    ms = e0.elapsed_ms(e1) / 20
    out[name] = {"items": int(mask.sum()), "ghost_cells": cells, "ms": ms, "algorithmic_GBs": 16e-9 * cells / (ms * 1e-3)}
print(json.dumps(out, indent=1))
