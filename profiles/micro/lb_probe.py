#!/usr/bin/env python
"""3 fused + 3 unfused LB steps on a 512x256x256 level in 64^3 boxes (for an ncu launch list)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402

sgc = load_package()
sgc.init(0)
n = [int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else [512, 256, 256, 64]
dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1)), (n[3],) * 3)
lb = sgc.LBLevel(dbl)
lb.advance(3, fuse=True)
lb.advance(3, fuse=False)
sgc.sync()
