"""Round-2 probe: stand-alone exchange (strip-sourced vs plain), the unfused drop-in step, and the
re-boxed step loop on small-box levels.  Usage: python profiles/micro/r02_probe.py [exchange|rebox|all]"""
import os
import sys
import json

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _loadpkg import load_package  # noqa: E402

sgc = load_package()
sgc.init(0)


def timed(fn, reps):
    e0, e1 = sgc.Event(), sgc.Event()
    sgc.sync()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    sgc.sync()
    return e0.elapsed_ms(e1) / reps


def level(n, box, per=7):
    dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
    a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
    rng = np.random.default_rng(1)
    u = rng.standard_normal(a.elems)
    a.upload(u)
    b.upload(u)
    cp = sgc.Copier().defineExchangeLD(a, per, 0)
    return dbl, a, b, cp


def exchange_probe(n=512, box=64):
    dbl, a, b, cp = level(n, box)
    out = {"level": "%d^3 in %d^3 boxes" % (n, box), "ghost_cells": cp.numCells(), "algorithmic_MB": 16e-6 * cp.numCells()}
    sgc.heat_sweep(b, a, 0.125)                   # b's outer strips are current now
    assert b.edgeStrips() == 2
    out["exchange_strips_ms"] = timed(lambda: b.exchange(cp), 20)
    os.environ["SGC_NO_XE"] = "1"
    out["exchange_plain_ms"] = timed(lambda: b.exchange(cp), 20)
    del os.environ["SGC_NO_XE"]
    out["sweep_ms"] = timed(lambda: sgc.heat_sweep(a, b, 0.125), 20)

    def step():
        b.exchange(cp)
        sgc.heat_sweep(a, b, 0.125)
        a.exchange(cp)
        sgc.heat_sweep(b, a, 0.125)
    t = timed(step, 10) / 2
    out["unfused_step_ms"] = t
    out["unfused_step_gcells"] = n ** 3 / (t * 1e-3) / 1e9
    for k in ("exchange_strips_ms", "exchange_plain_ms"):
        out[k.replace("_ms", "_GBs")] = out["algorithmic_MB"] / out[k]
    return out


def rebox_probe(n, box, nstep=100):
    dbl, a, b, cp = level(n, box)
    out = {"level": "%d^3 in %d^3 boxes" % (n, box), "nstep": nstep}
    for tag, env in (("rebox", None), ("per_box", "SGC_NO_REBOX")):
        if env:
            os.environ[env] = "1"
        cpx = sgc.Copier().defineExchangeLD(a, 7, 0)
        sgc.heat_run(a, b, cpx, 0.125, nstep)
        n0 = sgc.launch_count()
        t = timed(lambda: sgc.heat_run(a, b, cpx, 0.125, nstep), 3)
        out[tag + "_ms"] = t
        out[tag + "_gcells"] = n ** 3 * nstep / (t * 1e-3) / 1e9
        out[tag + "_launches_per_run"] = (sgc.launch_count() - n0) // 3
        out[tag + "_factors"] = cpx.reboxFactors()[1]
        if env:
            del os.environ[env]
    return out


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = {}
    if what in ("exchange", "all"):
        res["exchange_C2"] = exchange_probe(512, 64)
        res["exchange_32"] = exchange_probe(512, 32)
    if what == "c3":
        res["rebox_C3"] = rebox_probe(512, 16)
    if what in ("rebox", "all"):
        res["rebox_C3"] = rebox_probe(512, 16)
        res["rebox_32"] = rebox_probe(512, 32)
        res["rebox_8"] = rebox_probe(256, 8)
    print(json.dumps(res, indent=1))


def float_probe(n=512, box=64):
    """one-step sweep on float levels (streaming float instance) against the double one"""
    out = {}
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1,) * 3), box)
        a, b = sgc.LevelData(dbl, 1, 1, dtype=dt), sgc.LevelData(dbl, 1, 1, dtype=dt)
        u = np.random.default_rng(1).standard_normal(a.elems).astype(dt)
        a.upload(u)
        b.upload(u)

        def two():
            sgc.heat_sweep(b, a, 0.125)
            sgc.heat_sweep(a, b, 0.125)
        two()
        t = timed(two, 10) / 2
        out[tag] = {"sweep_ms": t, "gcells": n ** 3 / (t * 1e-3) / 1e9, "algorithmic_GBs": 2 * np.dtype(dt).itemsize * n ** 3 / (t * 1e-3) / 1e9}
    return out


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "float":
    print(json.dumps({"float_C2": float_probe(512, 64), "float_32": float_probe(512, 32)}, indent=1))
