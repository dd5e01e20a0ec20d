#!/usr/bin/env python
"""Lattice-Boltzmann application (D3Q19, 19-component level): device step loop vs the reference
application on the host cores.  Prints one JSON line per configuration.
  python profiles/micro/lb_bench.py [--cpu]
Algorithmic traffic per cell update of the fused step: 19 reads + 19 writes + 4 macroscopic writes
= 336 bytes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402

ALGO_BYTES = (19 + 19 + 4) * 8


def main():
    sgc = load_package()
    sgc.init(0)
    peak = 6460.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    out = []
    for (nx, ny, nz, bs, nstep) in ((64, 32, 32, 16, 400), (256, 128, 128, 32, 100), (512, 256, 256, 64, 40)):
        dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (nx - 1, ny - 1, nz - 1)), (bs, bs, bs))
        for fuse in (False, True):
            lb = sgc.LBLevel(dbl)
            lb.advance(5, fuse=fuse)
            sgc.sync()
            l0 = sgc.launch_count()
            e0, e1 = sgc.Event(), sgc.Event()
            e0.record()
            lb.advance(nstep, fuse=fuse)
            e1.record()
            ms = e0.elapsed_ms(e1)
            cells = nx * ny * nz
            rec = {"domain": [nx, ny, nz], "box": bs, "nstep": nstep, "fused": fuse, "ms_per_step": ms / nstep,
                   "mcell_updates_per_s": cells * nstep / (ms * 1e-3) / 1e6,
                   "algorithmic_GBs": cells * ALGO_BYTES * nstep / (ms * 1e-3) / 1e9,
                   "frac_of_copy_peak": cells * ALGO_BYTES * nstep / (ms * 1e-3) / 1e9 / peak,
                   "launches_per_step": (sgc.launch_count() - l0) / nstep}
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del lb
    if "--cpu" in sys.argv:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as po
        os.makedirs("plot", exist_ok=True)
        O = po.Oracle()
        for variant in ("", "_omp"):
            if not po.have_reference("_lb" + variant):
                continue
            R = po.ReferenceLB(variant)
            dlo, dhi, mb = (0, 0, 0), (63, 31, 31), (16, 16, 16)
            boxes, _, _ = O.layout(dlo, dhi, mb)
            nfi, nm = O.level_size(boxes, 1, 19), O.level_size(boxes, 1, 4)
            R.run(dlo, dhi, mb, 2, nfi, nm)
            t = time.time()
            nstep = 40
            R.run(dlo, dhi, mb, nstep, nfi, nm)
            dt = time.time() - t
            rec = {"cpu_reference": "latticeBoltzmann" + variant, "domain": [64, 32, 32], "box": 16, "nstep": nstep,
                   "ms_per_step": dt / nstep * 1e3, "mcell_updates_per_s": 64 * 32 * 32 * nstep / dt / 1e6,
                   "threads": os.cpu_count() if variant == "_omp" else 1}
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
