#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the REAL reference library
(oracle/_ref/libsgcref*.so, compiled from /root/reference by oracle/build_oracle.sh).

Run in the build container (the reference is not present on the GPU box):
    python oracle/gen_golden.py
The fixtures are small and committed; tests/test_oracle_golden.py checks the oracle against them
and the -m gpu tests check the CUDA path against them.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pyoracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (name, dlo, dhi, maxbox, ng, periodic, trim)
PLAN_CASES = [
    ("p222_per", (0, 0, 0), (7, 7, 7), (4, 4, 4), 1, 7, 0),
    ("p222_nonper", (0, 0, 0), (7, 7, 7), (4, 4, 4), 1, 0, 0),
    ("p333_off", (1, -1, 1), (12, 10, 12), (4, 4, 4), 1, 0, 0),       # testLayoutIterator.cpp domain
    ("p333_off_per", (1, -1, 1), (12, 10, 12), (4, 4, 4), 1, 7, 0),
    ("p444_trimcorner", (0, 0, 0), (15, 15, 15), (4, 4, 4), 1, 7, 8),
    ("p444_trimedgecorner", (0, 0, 0), (15, 15, 15), (4, 4, 4), 1, 7, 12),
    ("p444_ng2_perxy", (0, 0, 0), (15, 15, 15), (4, 4, 4), 2, 3, 0),
    ("p421_per", (0, 0, 0), (15, 7, 3), (4, 4, 4), 1, 7, 0),           # 1 box in z: self-periodic
    ("p_lb", (0, 0, 0), (63, 31, 31), (16, 16, 16), 1, 3, 8),          # latticeBoltzmann.cpp:16,36 layout
    ("p_aniso", (0, 0, 0), (11, 7, 9), (6, 4, 5), 1, 5, 0),
]

# (name, dlo, dhi, maxbox, ng, ncomp, periodic, trim)
EXCHANGE_CASES = [
    ("x222_c2", (0, 0, 0), (7, 7, 7), (4, 4, 4), 1, 2, 0, 0),           # testLevelData.cpp:126-171 shape
    ("x222_c2_per", (0, 0, 0), (7, 7, 7), (4, 4, 4), 1, 2, 7, 0),
    ("x333_off_per", (1, -1, 1), (12, 10, 12), (4, 4, 4), 1, 1, 7, 0),
    ("x_ng2", (0, 0, 0), (15, 15, 15), (8, 8, 8), 2, 1, 3, 0),
    ("x_trim", (0, 0, 0), (15, 15, 15), (4, 4, 4), 1, 3, 7, 8),
    ("x421", (0, 0, 0), (15, 7, 3), (4, 4, 4), 1, 1, 7, 0),
    ("x_aniso", (0, 0, 0), (11, 7, 9), (6, 4, 5), 1, 2, 5, 0),
]

# (name, dlo, dhi, maxbox, periodic, f, nstep, variant)
HEAT_CASES = [
    ("h16_b8_per_dyadic", (0, 0, 0), (15, 15, 15), (8, 8, 8), 7, 0.125, 10, ""),
    ("h16_b8_per_sixth_fma", (0, 0, 0), (15, 15, 15), (8, 8, 8), 7, 1.0 / 6.0, 10, ""),
    ("h16_b8_per_sixth_nofma", (0, 0, 0), (15, 15, 15), (8, 8, 8), 7, 1.0 / 6.0, 10, "_nofma"),
    ("h16_b4_dirichlet", (0, 0, 0), (15, 15, 15), (4, 4, 4), 0, 0.125, 7, ""),
    ("h_aniso_perxz", (2, 0, -3), (13, 7, 6), (6, 4, 5), 5, 0.1, 5, ""),
]

# known-answer hashes at larger sizes (only hashes are stored)
HEAT_HASH_CASES = [
    ("k64_b16", (0, 0, 0), (63, 63, 63), (16, 16, 16), 7, 0.125, 100, ""),
    ("k64_b32", (0, 0, 0), (63, 63, 63), (32, 32, 32), 7, 0.125, 100, ""),
    ("k64_b64", (0, 0, 0), (63, 63, 63), (64, 64, 64), 7, 0.125, 100, ""),
    ("k64_b16_sixth_fma", (0, 0, 0), (63, 63, 63), (16, 16, 16), 7, 1.0 / 6.0, 100, ""),
    ("k64_b16_sixth_nofma", (0, 0, 0), (63, 63, 63), (16, 16, 16), 7, 1.0 / 6.0, 100, "_nofma"),
    ("k128_b64_nonper", (0, 0, 0), (127, 127, 127), (64, 64, 64), 0, 0.125, 20, ""),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    O = po.Oracle()
    refs = {v: po.Reference(v) for v in ("", "_nofma")}
    R = refs[""]
    arrays = {}
    meta = {"plan": [], "exchange": [], "heat": [], "heat_hash": [], "laplacian": [], "wave": [], "fab": []}
    rng = np.random.default_rng(20261019)

    for name, dlo, dhi, mb, ng, per, trim in PLAN_CASES:
        for nproc in (1, 2):
            boxes, owner, nb = O.layout(dlo, dhi, mb)
            if len(boxes) % nproc:
                continue
            for p in range(nproc):
                key = "plan_%s_n%d_r%d" % (name, nproc, p)
                arrays[key] = R.copier_items(dlo, dhi, mb, ng, 1, per, trim, nproc, p)
                meta["plan"].append(dict(key=key, dlo=dlo, dhi=dhi, maxbox=mb, ng=ng, periodic=per, trim=trim,
                                         nproc=nproc, rank=p, nitem=int(len(arrays[key]))))
        b2, o2 = R.layout(dlo, dhi, mb, 1)
        arrays["layout_%s" % name] = b2

    for name, dlo, dhi, mb, ng, nc, per, trim in EXCHANGE_CASES:
        boxes, owner, nb = O.layout(dlo, dhi, mb)
        n = O.level_size(boxes, ng, nc)
        before = rng.integers(-1000, 1000, n).astype(np.float64) + rng.integers(0, 8, n) / 8.0
        after = before.copy()
        nitem = R.level_exchange(dlo, dhi, mb, ng, nc, per, trim, after, 0)
        arrays["xin_%s" % name] = before
        arrays["xout_%s" % name] = after
        meta["exchange"].append(dict(key=name, dlo=dlo, dhi=dhi, maxbox=mb, ng=ng, ncomp=nc, periodic=per,
                                     trim=trim, nitem=int(nitem)))

    def heat(dlo, dhi, mb, per, f, nstep, variant):
        boxes, owner, nb = O.layout(dlo, dhi, mb)
        u = O.fill_hash(boxes, 1, 0.0)
        nitem, _ = refs[variant].level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, u)
        g = O.gather(dlo, dhi, boxes, 1, 1, 0, u)
        return u, g, nitem

    for name, dlo, dhi, mb, per, f, nstep, variant in HEAT_CASES:
        u, g, nitem = heat(dlo, dhi, mb, per, f, nstep, variant)
        arrays["heat_%s" % name] = u
        meta["heat"].append(dict(key=name, dlo=dlo, dhi=dhi, maxbox=mb, periodic=per, f=f, nstep=nstep,
                                 contract=0 if variant == "_nofma" else 1, fnv=("%016x" % O.fnv1a64(g))))
    for name, dlo, dhi, mb, per, f, nstep, variant in HEAT_HASH_CASES:
        u, g, nitem = heat(dlo, dhi, mb, per, f, nstep, variant)
        meta["heat_hash"].append(dict(key=name, dlo=dlo, dhi=dhi, maxbox=mb, periodic=per, f=f, nstep=nstep,
                                      contract=0 if variant == "_nofma" else 1, fnv=("%016x" % O.fnv1a64(g)),
                                      fnv_level=("%016x" % O.fnv1a64(u)), nitem=int(nitem)))

    # laplacian driver (laplacian.cpp): linear field (B == 0) and the hash field, n = 16 and 24
    for n in (16, 24):
        ii = np.arange(1, n + 1, dtype=np.float64)
        A_lin = (ii[None, None, :] + ii[None, :, None] + ii[:, None, None]).ravel().copy()
        boxes = np.array([[1, 1, 1, n, n, n]], np.int32)
        A_hash = O.fill_hash(boxes, 0)
        for tag, A in (("lin", A_lin), ("hash", A_hash)):
            B, _ = R.laplacian(n, 16, A)
            B2, _ = R.laplacian(n, 16, A, boxiter=True)
            assert np.array_equal(B, B2)
            key = "lap_%s_n%d" % (tag, n)
            arrays[key] = B
            meta["laplacian"].append(dict(key=key, n=n, niter=16, init=tag, fnv=("%016x" % O.fnv1a64(B))))

    # wave (scalar form) 12^3, 8 steps, factor = cfl^2/3 with cfl = 0.5 (dyadic? no: tolerance/fma case)
    for name, factor, variant in (("wave_dyadic", 0.0625, ""), ("wave_fma", 0.25 / 3.0, ""), ("wave_nofma", 0.25 / 3.0, "_nofma")):
        dlo, dhi = (0, 0, 0), (11, 11, 11)
        boxes = np.array([[0, 0, 0, 11, 11, 11]], np.int32)
        u0 = O.fill_hash(boxes, 1, 0.0)
        u1 = u0.copy()
        refs[variant].wave(dlo, dhi, factor, 8, u0, u1)
        arrays[name + "_u0"] = u0
        arrays[name + "_u1"] = u1
        meta["wave"].append(dict(key=name, dlo=dlo, dhi=dhi, factor=factor, nstep=8,
                                 contract=0 if variant == "_nofma" else 1))

    # the SHIPPED wave time stepper (WavePatch::advance, USE_VEX pencil path) — valid cells are the contract
    if po.have_reference("_wave"):
        W = po.ReferenceWave()
        meta["wavepatch"] = []
        for name, h, nstep, init in (("wp16_hash", 16, 8, "hash"), ("wp13_hash", 13, 5, "hash"), ("wp16_pulse", 16, 10, "pulse")):
            if init == "hash":
                boxes = np.array([[0, 0, 0, h - 1, h - 1, h - 1]], np.int32)
                u_init = O.fill_hash(boxes, 1, 0.0)
                um1_init = u_init + rng.integers(0, 8, u_init.size) / 64.0
            else:
                u_init, um1_init, _ = W.run(h, 0.125, 0)          # WavePatch::initialData (sin^6 pulse)
            un, unm1, factor = W.run(h, 0.125, nstep, u_init, um1_init)
            arrays[name + "_u_init"], arrays[name + "_um1_init"] = u_init, um1_init
            arrays[name + "_un"], arrays[name + "_unm1"] = un, unm1
            meta["wavepatch"].append(dict(key=name, h=h, nstep=nstep, cfl=0.125, factor=factor, init=init))

    # the lattice-Boltzmann application (LBLevel::advance), from the reference built with and without
    # contraction; fi perturbed (ghosts included) so that every term of the collision is exercised
    if po.have_reference("_lb") and po.have_reference("_lb_nofma"):
        if not os.path.isdir("plot"):
            os.makedirs("plot")                      # LBLevel::writePlotFile's path prefix (never written: CGNS stand-in)
        LB = {"": po.ReferenceLB(""), "_nofma": po.ReferenceLB("_nofma")}
        meta["lb"], meta["lb_hash"], meta["plot"] = [], [], []
        for name, dlo, dhi, mb, nstep in (("lb_8x4x8_b4", (0, 0, 0), (7, 3, 7), (4, 4, 4), 5),
                                          ("lb_off_1layer", (2, -1, 0), (9, 6, 3), (4, 4, 4), 4)):
            boxes, owner, nb = O.layout(dlo, dhi, mb)
            fi0, m0 = O.lb_initial(boxes)
            fi0 = fi0 * (1 + rng.integers(-64, 64, fi0.size) / 4096.0)
            O.lb_macroscopic(boxes, fi0, m0)
            arrays[name + "_fi0"], arrays[name + "_macro0"] = fi0, m0
            for variant in ("", "_nofma"):
                fi, mac, mass = LB[variant].run(dlo, dhi, mb, nstep, fi0.size, m0.size, fi0, m0)
                if variant == "_nofma":
                    arrays[name + "_fi_nofma"] = fi
                arrays[name + "_macro" + variant] = mac
            meta["lb"].append(dict(key=name, dlo=dlo, dhi=dhi, maxbox=mb, nstep=nstep))
        # the shipped configuration (latticeBoltzmann.cpp:16,36): 64x32x32 in 16^3 boxes from initialData
        dlo, dhi, mb = (0, 0, 0), (63, 31, 31), (16, 16, 16)
        boxes, owner, nb = O.layout(dlo, dhi, mb)
        nfi, nm = O.level_size(boxes, 1, 19), O.level_size(boxes, 1, 4)
        for nstep in (1, 40):
            rec = dict(key="lb_shipped_%d" % nstep, dlo=dlo, dhi=dhi, maxbox=mb, nstep=nstep)
            for variant in ("", "_nofma"):
                fi, mac, mass = LB[variant].run(dlo, dhi, mb, nstep, nfi, nm)
                rec["fnv_macro" + variant] = "%016x" % O.fnv1a64(O.level_valid_out(boxes, 1, 4, mac))
                rec["fnv_fi" + variant] = "%016x" % O.fnv1a64(fi)
                rec["mass" + variant] = mass
            meta["lb_hash"].append(rec)
        # what LBLevel::writePlotFile hands to CGNS after 3 steps of a small level
        dlo, dhi, mb = (0, 0, 0), (7, 3, 7), (4, 4, 4)
        data, pm, names = LB["_nofma"].plot(dlo, dhi, mb, 3)
        arrays["plot_data"], arrays["plot_meta"] = data, pm
        meta["plot"].append(dict(key="plot", dlo=dlo, dhi=dhi, maxbox=mb, nstep=3, names=names))

    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    np.savez_compressed(os.path.join(OUT, "golden.npz"), **arrays)
    sz = os.path.getsize(os.path.join(OUT, "golden.npz"))
    print("wrote %d arrays, %.1f KiB" % (len(arrays), sz / 1024.0))
    for k in meta["heat_hash"]:
        print(k["key"], k["fnv"])


if __name__ == "__main__":
    main()
