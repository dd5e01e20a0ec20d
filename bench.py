#!/usr/bin/env python
"""bench.py — headline benchmark of the stencil-sweep + ghost-exchange path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C3|C4|C5]

Workloads (BASELINE.json configs):
  C2 (default at N = 1; configs[1]) : 512^3 level in 64^3 boxes, 1 ghost, periodic, double, hash-initialised;
                                      ONE STEP = 100 x { ghost exchange ; 7-point heat/Jacobi sweep }, ping-pong levels.
  C5 (default at N > 1; configs[4]) : weak scaling of C2: 512 x 512 x 512*N cut into z-slabs by the reference's
                                      ownership rule, z halo read over NVLink peer memory (NCCL schedule as fallback).
  C3 (configs[2])                   : the same level in 16^3 boxes (32 768 boxes, 851 968 motion items).
  C4 (configs[3])                   : 1024^3 heat equation over N GPUs (strong scaling; quoted at N = 8).
  C1 (configs[0])                   : the reference's laplacian driver: one 128^3 box, 16 x 3 directional passes.

One JSON line on rank 0: metric = Gcell-updates/s (whole job) + `roofline` (dominant kernel against the measured
HBM copy bandwidth, 16 algorithmic bytes per cell-update; `dram_*` keys give the fraction in real DRAM bytes),
`parity` (the device result of `cpu_sweeps` sweeps on the FULL-SIZE level against the reference's own CPU run of
the same level, bit for bit; at N > 1 `parity_nranks`: slab-partitioned levels against the reference's global run),
`cpu_baseline` (the reference's CPU code, oracle/_ref, timed on this host on a bounded sample), `e2e` (the same
job through the C ABI with HOST buffers: H2D + 100 sweeps + D2H per step), `unfused` (the drop-in path of an
unmodified application: exchange(); sweep() per step), `configs` (N = 1 default run: C1 and C3 measured in the
same run), `clocks`, `gpu_launches`.

`--impl reference` times the reference's CPU implementation (oracle/_ref/libsgcref*.so, compiled from
/root/reference in the build container; falls back to the oracle port) on the same config — at N > 1 the same
GLOBAL level, with OMP_NUM_THREADS forced to the host's core count (torchrun exports 1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_CELL = 16.0      # read u once + write unew once, double
LAP_BYTES_PER_CELL = 24.0       # laplacian driver: read A, read B, write B per cell and iteration
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=["C1", "C2", "C3", "C4", "C5"],
                    help="BASELINE config (default: C2 at N = 1, C5 = its weak-scaling form at N > 1)")
    ap.add_argument("--sweeps", type=int, default=100, help="sweeps per step (config: 100)")
    ap.add_argument("--domain", type=int, default=0, help="cells per side per GPU (x, y and z-per-GPU)")
    ap.add_argument("--domain-z", type=int, default=0, help="z cells per GPU if different from --domain")
    ap.add_argument("--box", type=int, default=0)
    ap.add_argument("--f", type=float, default=0.125)
    ap.add_argument("--e2e-steps", type=int, default=2, help="single-batch end-to-end repetitions (0: no e2e at all)")
    ap.add_argument("--e2e-batches", type=int, default=32, help="batches streamed through the pipelined end-to-end leg")
    ap.add_argument("--cpu-sweeps", type=int, default=12, help="sweeps of the CPU baseline sample / the parity run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 sub-measurements of the default run")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.config is None:
        a.config = "C2" if world == 1 else "C5"
    if a.config == "C3":
        a.domain, a.box = a.domain or 512, a.box or 16
    elif a.config == "C4":
        a.domain, a.box = a.domain or 1024, a.box or 64
        if not a.domain_z:
            if a.domain % (world * a.box):
                raise SystemExit("C4: %d ranks do not divide the %d box layers" % (world, a.domain // a.box))
            a.domain_z = a.domain // world
    else:
        a.domain, a.box = a.domain or 512, a.box or 64
    return a


def hash_field(lo, n):
    """SURVEY §8c seedless dyadic init on the box [lo, lo+n): values k/1024."""
    i = (np.arange(lo[0], lo[0] + n[0], dtype=np.uint32) * np.uint32(73856093))[None, None, :]
    j = (np.arange(lo[1], lo[1] + n[1], dtype=np.uint32) * np.uint32(19349663))[None, :, None]
    k = (np.arange(lo[2], lo[2] + n[2], dtype=np.uint32) * np.uint32(83492791))[:, None, None]
    return ((i ^ j ^ k) % np.uint32(1024)).astype(np.float64) / 1024.0


def fill_level_host(host, boxes, ng):
    off = 0
    for b in boxes:
        n = b[3:] - b[:3] + 1
        m = n + 2 * ng
        f = host[off:off + int(np.prod(m))].reshape(m[2], m[1], m[0])
        f[...] = 0.0
        f[ng:ng + n[2], ng:ng + n[1], ng:ng + n[0]] = hash_field(b[:3], n)
        off += int(np.prod(m))
    return host


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 8] or [r for (_, r) in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def load_traffic(key="sweep_dram_bytes_per_launch"):
    """DRAM bytes per launch of a kernel from the last committed `ncu --set full` capture (profiles/traffic.json):
    a constant of the kernel on this level shape, NOT measured by this run."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ------------------------------------------------------------------------------ the reference's CPU code
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    return pyoracle


def cpu_reference_run(dom, box, f, sweeps, variants=("", "_omp"), periodic=7, keep_field=False):
    """Time the reference's CPU implementation (kind "reference": oracle/_ref built from /root/reference; else
    kind "port": the oracle) on `sweeps` x {exchange; sweep} of the level `dom` = (nx, ny, nz) in box^3 boxes.
    Returns a list of dicts (value Gcell/s, cores, kind, ms split[, field])."""
    po = _oracle_module()
    O = po.Oracle()
    dlo, dhi, mb = (0, 0, 0), (dom[0] - 1, dom[1] - 1, dom[2] - 1), (box,) * 3
    boxes, _, _ = O.layout(dlo, dhi, mb)
    u0 = np.empty(O.level_size(boxes, 1, 1))
    fill_level_host(u0, boxes, 1)
    cells = float(dom[0]) * dom[1] * dom[2]
    runs = []
    if po.have_reference(""):
        for variant in variants:
            if not po.have_reference(variant):
                continue
            R = po.Reference(variant)
            cores = R.set_threads(host_cores()) if variant == "_omp" else 1
            u = u0.copy()
            t0 = time.time()
            _, tm = R.level_heat(dlo, dhi, mb, 1, periodic, 0, f, sweeps, u)
            wall = time.time() - t0
            ms = float(tm[1] + tm[2])
            runs.append(dict(kind="reference", variant=variant or "serial", cores=cores,
                             ms=ms, exchange_ms=float(tm[1]), sweep_ms=float(tm[2]), plan_ms=float(tm[0]), wall_s=wall,
                             value=cells * sweeps / (ms * 1e-3) / 1e9))
            if keep_field:
                runs[-1]["field"] = u
    else:
        u = u0.copy()
        _, tm = O.level_heat(dlo, dhi, mb, 1, periodic, 0, f, sweeps, u, 1)
        ms = float(tm[1] + tm[2])
        runs.append(dict(kind="port", variant="oracle", cores=1, ms=ms, exchange_ms=float(tm[1]), sweep_ms=float(tm[2]),
                         plan_ms=float(tm[0]), wall_s=0.0, value=cells * sweeps / (ms * 1e-3) / 1e9))
        if keep_field:
            runs[-1]["field"] = u
    return runs


def cpu_laplacian_run(n, iters, A):
    """The reference's laplacian driver loop (laplacian.cpp:90-110) on a box of n^3: (B, ms, kind)."""
    po = _oracle_module()
    if po.have_reference(""):
        B, ms = po.Reference("").laplacian(n, iters, A)
        return B, float(ms), "reference"
    O = po.Oracle()
    t0 = time.time()
    B = O.laplacian(n, iters, A)
    return B, (time.time() - t0) * 1e3, "port"


def strip_fields(runs):
    return [{k: v for k, v in r.items() if k != "field"} for r in runs]


def global_domain(a, n):
    return (a.domain, a.domain, (a.domain_z or a.domain) * n)


def workload_config(a, n):
    if a.config == "C1":
        return {"workload": "C1: reference laplacian driver, one 128^3 box A on (1..128)^3, B on (2..127)^3, "
                            "16 x 3 directional passes B -= 0.5*T_d(A) per step",
                "box": 128, "iterations_per_step": 16, "ncomp": 1,
                "l2_policy": "33 MB working set: L2-resident by construction of the reference's own driver"}
    zper = a.domain_z or a.domain
    nbox = (a.domain // a.box) ** 2 * (zper // a.box)
    return {"workload": "%s: %dx%dx%d level in %d^3 boxes, 1 ghost, periodic xyz, %d x {exchange; 7-pt heat sweep} per step"
                        % (a.config, a.domain, a.domain, zper * n, a.box, a.sweeps),
            "domain": [a.domain, a.domain, zper * n], "box": a.box, "nghost": 1, "ncomp": 1,
            "sweeps_per_step": a.sweeps, "f": a.f, "periodic": "xyz", "partition": "z-slabs, owner = idx/(nbox/N)",
            "l2_policy": "inputs larger than L2 (2 x %.2f GB per GPU)" % (nbox * (a.box + 2) ** 3 * 8 / 1e9)}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["OMP_NUM_THREADS"] = str(host_cores())      # torchrun exports 1; the shim also sets it explicitly
    if a.config == "C1":
        return run_reference_arm_c1(a)
    dom = global_domain(a, world)
    # bounded sample: each step = `sample_sweeps` sweeps of the SAME global level our arm runs
    sample_sweeps = max(1, min(max(2, a.cpu_sweeps // world), a.sweeps))
    probe = cpu_reference_run(dom, a.box, a.f, 1)
    best = max(probe, key=lambda r: r["value"])
    variants = ("_omp",) if best["variant"] == "_omp" else ("",)
    vals = []
    for s in range(a.warmup + a.steps):
        r = cpu_reference_run(dom, a.box, a.f, sample_sweeps, variants=variants)[0]
        if s >= a.warmup:
            vals.append(r)
    split = vals[-1]
    ms = float(np.mean([r["ms"] for r in vals])) * (a.sweeps / sample_sweeps)
    cells = float(dom[0]) * dom[1] * dom[2]
    value = cells * a.sweeps / (ms * 1e-3) / 1e9
    sample = "%d of %d sweeps per step on the full %dx%dx%d level in %d^3 boxes (exchange+sweep, Stopwatch)" % (
        sample_sweeps, a.sweeps, dom[0], dom[1], dom[2], a.box)
    line = {
        "impl": "reference", "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if a.config == "C4" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, world),
        "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": split["cores"], "kind": split["kind"],
                         "variant": split["variant"], "sample": sample, "exchange_ms_per_sweep": split["exchange_ms"] / sample_sweeps,
                         "sweep_ms_per_sweep": split["sweep_ms"] / sample_sweeps, "host_cores_available": host_cores(),
                         "probe": strip_fields(probe)},
        "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_reference_arm_c1(a):
    n, iters = 128, 16
    A = (np.add.outer(np.add.outer(np.arange(1, n + 1.0), np.arange(1, n + 1.0)), np.arange(1, n + 1.0))).ravel()
    ms_all = []
    for s in range(a.warmup + a.steps):
        _, ms, kind = cpu_laplacian_run(n, iters, A)
        if s >= a.warmup:
            ms_all.append(ms)
    ms = float(np.mean(ms_all))
    value = (n - 2.0) ** 3 * iters / (ms * 1e-3) / 1e9
    emit({"impl": "reference", "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": a.gpus,
          "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, 1),
          "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": 1, "kind": kind,
                           "sample": "the whole step (16 iterations of the driver loop, laplacian.cpp:90-110)"},
          "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


# ------------------------------------------------------------------------------ our arm
_REAL_STDOUT = None
_FULL_AFFINITY = None


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank (and therefore the pinned host buffers it allocates, first touch) to the CPUs of the NUMA node
    its GPU hangs off — what `numactl --cpunodebind --membind` does for a production launch.  Returns a description
    for the JSON line; silently does nothing where sysfs / nvidia-smi do not tell (single-node hosts)."""
    global _FULL_AFFINITY
    try:
        _FULL_AFFINITY = os.sched_getaffinity(0)
        bus = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None, "note": "platform reports no NUMA affinity for the GPU"}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= _FULL_AFFINITY
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus_bound": len(cpus), "pci_bus_id": bus}
    except Exception as ex:
        return {"numa_node": None, "note": repr(ex)}


def restore_full_affinity():
    if _FULL_AFFINITY:
        try:
            os.sched_setaffinity(0, _FULL_AFFINITY)
        except OSError:
            pass


def emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints (NCCL's version
    banner, torchrun notices) was redirected to stderr by quiet_stdout()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def kernel_roofline(sgc, cells_local, sweeps_done, prof, peak, peak_src, box):
    """Roofline of the dominant kernel from the library's per-launch event pairs (sgc_profile_*): inside
    sgc_heat_run that is the two-step sweep (k_heat_t2, class 2: ONE launch = TWO cell-updates per cell = 2 x 16
    algorithmic bytes per cell) wherever the level (or its re-boxed form) is covered by it; the remaining steps and
    other levels use the one-step sweep (class 0).  `frac` > 1 means the kernel moves fewer DRAM bytes than the
    algorithmic figure: `dram_frac_of_peak` (ncu traffic) and `dram_frac_of_compulsory` (read the level once +
    write it once per launch) say how close the launch is to the HBM bound in real bytes."""
    (sweep_n, sweep_ms), (exch_n, exch_ms), (t2_n, t2_ms) = prof
    sweep_ms_per_sweep = (sweep_ms + t2_ms) / max(sweeps_done, 1)
    one_step = {"launches_timed": sweep_n, "avg_launch_ms": sweep_ms / max(sweep_n, 1),
                "achieved_GBs": (cells_local * ALGO_BYTES_PER_CELL * (sweeps_done - 2 * t2_n) / (sweep_ms * 1e-3) / 1e9
                                 if sweep_ms > 0 else 0.0),
                "traffic": load_traffic("sweep_dram_bytes_per_launch") if box == 64 else None}
    if t2_n > 0:
        kernel, n_l, ms_l, upd = "two-step heat sweep k_heat_t2 (class 2)", t2_n, t2_ms, 2
        traffic = load_traffic("t2_dram_bytes_per_launch")
    else:
        kernel, n_l, ms_l, upd = "heat sweep k_heat_stream (class 0)", sweep_n, sweep_ms, 1
        traffic = one_step["traffic"]
    # with N > 1 and the NCCL schedule a step issues 2-3 one-step launches (interior + boundary ranges):
    # there use bytes per sweep rather than per launch
    if t2_n == 0 and sweep_n != sweeps_done:
        avg_ms = sweep_ms / max(sweeps_done, 1)
    else:
        avg_ms = ms_l / max(n_l, 1)
    bytes_per_launch = cells_local * ALGO_BYTES_PER_CELL * upd
    compulsory = cells_local * 16.0          # the launch reads every cell once and writes every cell once
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
            "traffic_source": "ncu --set full capture committed under profiles/ (traffic.json), per launch on the "
                              "512^3/64^3 fab shape; not re-measured by this run",
            "algorithmic_bytes_per_launch": bytes_per_launch, "cell_updates_per_cell_per_launch": upd,
            "avg_launch_ms": avg_ms, "launches_timed": n_l,
            "frac_of_nominal_8TBs": achieved / 8000.0,
            "dram_frac_of_peak": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic and avg_ms > 0 else None,
            "dram_frac_of_compulsory": (compulsory / (avg_ms * 1e-3) / 1e9 / peak) if avg_ms > 0 else None,
            "one_step_kernel": one_step,
            "sweep_only_gcells": cells_local / (sweep_ms_per_sweep * 1e-3) / 1e9 if sweep_ms_per_sweep > 0 else 0.0,
            "copy_launches": exch_n, "copy_ms_per_sweep": exch_ms / max(sweeps_done, 1),
            "sweep_ms_per_sweep": sweep_ms_per_sweep}


def read_profile(sgc):
    return sgc.profile_read(0), sgc.profile_read(1), sgc.profile_read(2)


def parity_full_size(sgc, A, B, cp, host_init, a, dom, box, variants=("", "_omp"), label=None):
    """`cpu_sweeps` sweeps of the FULL-SIZE level from the hash field on the device, against the reference's own
    CPU run of the same level (every cell of the result level, ghost cells included).  Returns (parity dict,
    cpu runs)."""
    po = _oracle_module()
    O = po.Oracle()
    runs = cpu_reference_run(dom, box, a.f, a.cpu_sweeps, variants=variants, keep_field=True)
    A.upload(host_init)
    B.upload(host_init)
    res = sgc.heat_run(A, B, cp, a.f, a.cpu_sweeps)
    got = res.download()
    ref = runs[-1]["field"]
    same = all(np.array_equal(r["field"], ref) for r in runs)
    ok = bool(np.array_equal(got, ref)) and same
    par = {"config": "%s full size: %dx%dx%d in %d^3 boxes" % (label or a.config, dom[0], dom[1], dom[2], box),
           "sweeps": a.cpu_sweeps, "against": runs[-1]["kind"], "variants_agree": bool(same),
           "cells_compared": int(got.size), "ghost_cells_included": True,
           "fnv": "%016x" % O.fnv1a64(got), "fnv_reference": "%016x" % O.fnv1a64(ref), "ok": ok}
    return par, runs


def parity_nranks(sgc, dist, world, rank):
    """Slab-partitioned levels through the multi-GPU step loop (peer-memory two-step passes, fused one-step sweeps,
    NCCL exchange) against the CPU reference's run of the GLOBAL level: every rank checks every cell of its boxes,
    ghost cells included (the cases of tests/multi_worker.py)."""
    import torch
    po = _oracle_module()
    O = po.Oracle()
    R = po.Reference("") if po.have_reference("") else None
    cases = (((0, 0, 0), (31, 31, 32 * world - 1), (16, 16, 16), 7, 0.125, 12),
             ((0, 0, 0), (63, 63, 64 * world - 1), (64, 64, 64), 7, 1.0 / 6.0, 6),
             ((0, 0, 0), (127, 63, 128 * world - 1), (64, 64, 64), 7, 0.2, 11),
             ((0, 0, 0), (127, 63, 128 * world - 1), (64, 64, 64), 3, 0.2, 10),
             ((0, 0, 0), (63, 127, 64 * world - 1), (64, 64, 64), 0, 0.125, 9),
             ((0, 0, 0), (63, 31, 64 * world - 1), (32, 32, 32), 3, 0.125, 9),
             ((0, 0, 0), (15, 15, 8 * world - 1), (8, 8, 8), 0, 0.125, 5),
             ((0, 0, 0), (11, 7, 10 * world - 1), (6, 4, 5), 3, 0.2, 4))
    ok, t2_total = True, 0
    for dlo, dhi, mb, per, f, nstep in cases:
        dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb, world, rank)
        allboxes, _, _ = O.layout(dlo, dhi, mb)
        u0 = O.fill_hash(allboxes, 1, 0.0)
        goff = O.level_offsets(allboxes, 1, 1)
        want = u0.copy()
        if R is not None:
            R.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want)
        else:
            O.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want, 1)
        b0, b1 = dbl.localIdxBegin(), dbl.localIdxEnd()
        end = goff[b1] if b1 < len(goff) else len(u0)
        mine = u0[goff[b0]:end].copy()
        la, lb = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
        la.upload(mine)
        lb.upload(mine)
        cpx = sgc.Copier().defineExchangeLD(la, per, 0)
        sgc.profile_enable(True)
        sgc.profile_read(2)
        res = sgc.heat_run(la, lb, cpx, f, nstep)
        n2, _ = sgc.profile_read(2)
        sgc.profile_enable(False)
        t2_total += n2
        ok = ok and bool(np.array_equal(res.download(), want[goff[b0]:end]))
    t = torch.tensor([1.0 if ok else 0.0, float(t2_total)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"cases": len(cases), "ranks": world, "against": "reference" if R is not None else "port",
            "ghost_cells_included": True, "two_step_passes_min_over_ranks": int(t[1]), "ok": bool(t[0] == 1.0)}


def measure_heat(sgc, A, B, cp, f, sweeps, steps, warmup):
    """`steps` x sgc_heat_run(sweeps) after `warmup`, device-timed; returns (ms per step, launches, profile)."""
    cur, oth = A, B

    def step():
        nonlocal cur, oth
        res = sgc.heat_run(cur, oth, cp, f, sweeps)
        if res is not cur:
            cur, oth = oth, cur

    for _ in range(warmup):
        step()
    sgc.sync()
    sgc.profile_enable(True)
    read_profile(sgc)
    l0 = sgc.launch_count()
    e0, e1 = sgc.Event(), sgc.Event()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sgc.sync()
    ms = e0.elapsed_ms(e1) / steps
    launches = sgc.launch_count() - l0
    prof = read_profile(sgc)
    sgc.profile_enable(False)
    return ms, launches, prof, cur


def sub_config_c3(sgc, a, peak, peak_src):
    """BASELINE configs[2] inside the default run: 32 768 boxes of 16^3 (512^3), the same 100-sweep step."""
    box = 16
    dom = (a.domain, a.domain, a.domain)
    dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (dom[0] - 1, dom[1] - 1, dom[2] - 1)), box)
    A, B = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
    host = np.empty(A.elems)
    fill_level_host(host, dbl.local_boxes(), 1)
    A.upload(host)
    B.upload(host)
    t0 = time.time()
    cp = sgc.Copier().defineExchangeLD(A, 7, 0)
    plan_ms = (time.time() - t0) * 1e3
    ms, launches, prof, _ = measure_heat(sgc, A, B, cp, a.f, a.sweeps, 3, 3)
    cells = float(dom[0]) * dom[1] * dom[2]
    out = {"workload": "C3: %d^3 level in 16^3 boxes (%d boxes), 1 ghost, periodic xyz, %d x {exchange; sweep} per step"
                       % (a.domain, A.size(), a.sweeps),
           "value": cells * a.sweeps / (ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "ms_per_step": ms, "steps": 3, "warmup": 3,
           "gpu_launches_per_step": int(launches // 3), "motion_items": cp.numMotionItem(), "plan_build_ms": plan_ms,
           "rebox_factors": list(cp.reboxFactors()[1]),
           "roofline": kernel_roofline(sgc, cells, 3 * a.sweeps, prof, peak, peak_src, 64)}
    if not a.no_cpu_baseline:
        try:
            # (the reference's OpenMP build collapses on the 851 968-item exchange: its serial build is the faster one)
            par, runs = parity_full_size(sgc, A, B, cp, host, a, dom, box, variants=("",), label="C3")
            out["parity"] = par
            best = max(runs, key=lambda r: r["value"])
            out["cpu_baseline"] = {"value": best["value"], "unit": "Gcell-updates/s", "cores": best["cores"], "kind": best["kind"],
                                   "variant": best["variant"], "sample": "%d sweeps of the full level" % a.cpu_sweeps,
                                   "exchange_ms_per_sweep": best["exchange_ms"] / a.cpu_sweeps,
                                   "sweep_ms_per_sweep": best["sweep_ms"] / a.cpu_sweeps}
        except Exception as ex:
            out["parity"] = {"ok": None, "error": repr(ex)}
    return out


def sub_config_c4(sgc, dist, a, world, rank, peak, peak_src):
    """BASELINE configs[3] inside the default 8-GPU run: the 1024^3 heat equation in 64^3 boxes over the 8 GPUs (128
    planes = two box layers per GPU: EVERY box reads a z neighbour over NVLink), the same 100-sweep step.  Collective:
    every rank calls it."""
    import torch
    n, box = 1024, 64
    dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (n - 1, n - 1, n - 1)), box, world, rank)
    A, B = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
    host = np.empty(A.elems)
    fill_level_host(host, dbl.local_boxes(), 1)
    A.upload(host)
    B.upload(host)
    cp = sgc.Copier().defineExchangeLD(A, 7, 0)
    sgc.sync()
    dist.barrier()
    ms, launches, prof, cur = measure_heat(sgc, A, B, cp, a.f, a.sweeps, 3, 2)
    out = cur.download()
    m = box + 2
    sums = torch.tensor([float(out.reshape(A.size(), m, m, m)[:, 1:-1, 1:-1, 1:-1].sum(dtype=np.float64)),
                         float(host.reshape(A.size(), m, m, m)[:, 1:-1, 1:-1, 1:-1].sum(dtype=np.float64))], dtype=torch.float64)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    t = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    cells_local = float(n) * n * (n // world)
    return {"workload": "C4: 1024^3 heat equation in 64^3 boxes over %d GPUs (z-slabs of %d planes), periodic, %d x {exchange; sweep} "
                        "per step, z halo read over NVLink peer memory inside the sweeps" % (world, n // world, a.sweeps),
            "value": float(n) ** 3 * a.sweeps / (ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "ms_per_step": ms, "steps": 3, "warmup": 2,
            "scaling": "strong", "n_gpus": world, "gpu_launches_per_step": int(launches // 3), "remote_items": cp.numRemoteItem(),
            "sum_conserved": bool(abs(float(sums[0]) - float(sums[1])) <= 1e-9 * abs(float(sums[1]))),
            "roofline": kernel_roofline(sgc, cells_local, 3 * a.sweeps, prof, peak, peak_src, 64)}


def run_c1(sgc, a, peak, peak_src, steps, warmup, with_cpu=True):
    """BASELINE configs[0]: the reference's laplacian driver (laplacian.cpp:11-120) with n = 128: A on (1..128)^3,
    B on (2..127)^3, 16 iterations of the three directional passes B -= 0.5 * T_d(A).  One step = the 16
    iterations, one launch each (the un-fused 24 B per cell and iteration of SURVEY §8d)."""
    n, iters = 128, 16
    la = sgc.LevelData(None, 1, 0, boxes=np.array([[1, 1, 1, n, n, n]], np.int32))
    lb = sgc.LevelData(None, 1, 0, boxes=np.array([[2, 2, 2, n - 1, n - 1, n - 1]], np.int32))
    idx = np.arange(1, n + 1, dtype=np.uint32)
    hashA = (((idx * np.uint32(73856093))[None, None, :] ^ (idx * np.uint32(19349663))[None, :, None] ^
              (idx * np.uint32(83492791))[:, None, None]) % np.uint32(1024)).astype(np.float64).ravel() / 1024.0
    linA = np.add.outer(np.add.outer(np.arange(1, n + 1.0), np.arange(1, n + 1.0)), np.arange(1, n + 1.0)).ravel()
    pinA, pinB = sgc.PinnedBuffer(la.elems), sgc.PinnedBuffer(lb.elems)
    pinA.array[:] = hashA
    la.upload(pinA.array)
    zero = np.zeros(lb.elems)

    def step():
        sgc.laplacian_apply(lb, la, 0.5, iters, 0)

    for _ in range(warmup):
        lb.upload(zero)
        step()
    sgc.sync()
    ms_all = []
    l0 = sgc.launch_count()
    for _ in range(steps):
        lb.upload(zero)
        sgc.sync()
        e0, e1 = sgc.Event(), sgc.Event()
        e0.record()
        step()
        e1.record()
        sgc.sync()
        ms_all.append(e0.elapsed_ms(e1))
    launches = sgc.launch_count() - l0
    ms = float(np.mean(ms_all))
    gotB = lb.download()
    # end to end through the C ABI with host buffers: upload A and a zero B, 16 iterations, download B
    e2e_ms = []
    for s in range(3):
        sgc.sync()
        e0, e1 = sgc.Event(), sgc.Event()
        e0.record()
        la.upload(pinA.array, wait=False)
        lb.upload(zero, wait=False)
        step()
        lb.download(out=pinB.array)
        e1.record()
        sgc.sync()
        if s:
            e2e_ms.append(e0.elapsed_ms(e1))
    # the same 16 iterations in ONE launch (A is constant over the iterations, so B's 48 updates per cell are applied
    # in the reference's order in registers: bit-identical, 24 B per cell for the whole step instead of per iteration)
    fused_ms = []
    for s in range(4):
        lb.upload(zero)
        sgc.sync()
        e0, e1 = sgc.Event(), sgc.Event()
        e0.record()
        sgc.laplacian_apply(lb, la, 0.5, iters, 1)
        e1.record()
        sgc.sync()
        if s:
            fused_ms.append(e0.elapsed_ms(e1))
    fused_same = bool(np.array_equal(lb.download(), gotB))
    cells = (n - 2.0) ** 3
    upd = cells * iters
    gb = upd * LAP_BYTES_PER_CELL / (ms * 1e-3) / 1e9
    out = {"workload": workload_config(argparse.Namespace(config="C1"), 1)["workload"],
           "value": upd / (ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
           "gpu_launches_per_step": int((launches - steps) // steps),      # minus the zero-fill conversion launch
           "roofline": {"bound": "hbm", "kernel": "k_sweep7<LapOp> (one launch per iteration)", "achieved": gb, "peak": peak,
                        "unit": "GB/s", "frac": gb / peak, "peak_source": peak_src, "traffic": None,
                        "algorithmic_bytes_per_launch": cells * LAP_BYTES_PER_CELL, "avg_launch_ms": ms / iters,
                        "note": "A (16.8 MB) and B (16.0 MB) stay in the 126 MB L2 between launches: the launch is bound "
                                "by launch latency and L2, the HBM figure is the contract's denominator only"},
           "fused_iterations": {"value": upd / (float(np.mean(fused_ms)) * 1e-3) / 1e9, "unit": "Gcell-updates/s",
                                "ms_per_step": float(np.mean(fused_ms)), "launches_per_step": 1, "same_bits_as_unfused": fused_same},
           "e2e": {"value": upd / (float(np.mean(e2e_ms)) * 1e-3) / 1e9, "unit": "Gcell-updates/s",
                   "h2d_bytes_per_step": int((la.elems + lb.elems) * 8), "d2h_bytes_per_step": int(lb.elems * 8),
                   "ms_per_step": float(np.mean(e2e_ms))}}
    if with_cpu:
        try:
            refB, cpu_ms, kind = cpu_laplacian_run(n, iters, hashA)
            la.upload(linA)
            lb.upload(zero)
            step()
            lin_zero = bool((lb.download() == 0.0).all())            # laplacian.cpp:80-86,115-120: B == 0 exactly
            out["parity"] = {"config": "C1 full size (n = 128), hash field", "against": kind,
                             "ok": bool(np.array_equal(gotB, refB)) and lin_zero, "linear_field_B_is_zero": lin_zero}
            out["cpu_baseline"] = {"value": upd / (cpu_ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "cores": 1, "kind": kind,
                                   "sample": "the whole step (16 iterations of the driver loop, laplacian.cpp:90-110)"}
        except Exception as ex:
            out["parity"] = {"ok": None, "error": repr(ex)}
    return out


def main():
    a = parse()
    quiet_stdout()
    if a.impl == "reference":
        return run_reference_arm(a)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    n_gpus = world

    numa = bind_to_gpu_numa_node(local_rank)
    from _loadpkg import load_package
    sgc = load_package()
    sgc.init(local_rank)
    peak, peak_src = measured_hbm_peak()

    if a.config == "C1":
        if rank == 0:
            sampler = ClockSampler(local_rank)
            sampler.start()
            tw0 = time.time()
            restore_full_affinity()
            c1 = run_c1(sgc, a, peak, peak_src, max(a.steps, 5), max(a.warmup, 3), with_cpu=not a.no_cpu_baseline)
            clocks = sampler.stop(tw0, time.time())
            name, sms, _ = sgc.device_info()
            line = {"metric": "Gcell-updates/s", "value": c1["value"], "unit": "Gcell-updates/s", "n_gpus": 1,
                    "steps": c1["steps"], "warmup": c1["warmup"], "ms_per_step": c1["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": workload_config(a, 1), "roofline": c1["roofline"], "cpu_baseline": c1.get("cpu_baseline"),
                    "e2e": c1["e2e"], "parity": c1.get("parity"), "clocks": clocks,
                    "gpu_launches": c1["gpu_launches_per_step"] * c1["steps"], "device": name, "sm_count": sms}
            emit(line)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    zper = a.domain_z or a.domain
    if world > 1:
        ids = [sgc.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sgc.comm_init(world, rank, ids[0])

    # multi-GPU parity inside the bench's own process group, before anything is timed
    par_n = parity_nranks(sgc, dist, world, rank) if world > 1 and not a.no_cpu_baseline else None

    dom = sgc.Box((0, 0, 0), (a.domain - 1, a.domain - 1, zper * world - 1))
    dbl = sgc.DisjointBoxLayout(dom, a.box, world, rank)
    A = sgc.LevelData(dbl, 1, 1)
    B = sgc.LevelData(dbl, 1, 1)
    cells_local = float(a.domain) ** 2 * zper
    cells_total = cells_local * world

    pin_in = sgc.PinnedBuffer(A.elems)
    pin_out = sgc.PinnedBuffer(A.elems)
    fill_level_host(pin_in.array, dbl.local_boxes(), 1)
    A.upload(pin_in.array)
    B.upload(pin_in.array)
    t0 = time.time()
    cp = sgc.Copier().defineExchangeLD(A, 7, 0)
    plan_ms = (time.time() - t0) * 1e3

    def barrier():
        sgc.sync()
        if dist is not None:
            dist.barrier()
        sgc.sync()

    cur, oth = A, B

    def step():
        nonlocal cur, oth
        res = sgc.heat_run(cur, oth, cp, a.f, a.sweeps)
        if res is not cur:
            cur, oth = oth, cur

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    sgc.profile_enable(True)
    read_profile(sgc)
    launches0 = sgc.launch_count()
    e0, e1 = sgc.Event(), sgc.Event()
    barrier()
    tw0 = time.time()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    tw1 = time.time()
    ms_total = e0.elapsed_ms(e1)
    launches = sgc.launch_count() - launches0
    prof = read_profile(sgc)
    sgc.profile_enable(False)
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None

    if dist is not None:
        import torch
        t = torch.tensor([ms_total], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t[0])
    ms_per_step = ms_total / a.steps
    value = cells_total * a.sweeps / (ms_per_step * 1e-3) / 1e9

    # conservation check of the device result (periodic heat equation conserves the sum)
    out = cur.download(out=pin_out.array)
    m = a.box + 2
    valid = out.reshape(A.size(), m, m, m)[:, 1:-1, 1:-1, 1:-1]
    mask_sum = float(valid.sum(dtype=np.float64))
    ref_sum = float(pin_in.array.reshape(A.size(), m, m, m)[:, 1:-1, 1:-1, 1:-1].sum(dtype=np.float64))
    if dist is not None:
        import torch
        t = torch.tensor([mask_sum, ref_sum], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        mask_sum, ref_sum = float(t[0]), float(t[1])
    conserved = abs(mask_sum - ref_sum) <= 1e-9 * abs(ref_sum)

    sweeps_done = a.steps * a.sweeps
    roofline = kernel_roofline(sgc, cells_local, sweeps_done, prof, peak, peak_src, a.box)
    roofline["exchange_bytes_per_sweep"] = 16.0 * cp.numCells()

    # The drop-in path of an unmodified application: LevelData::exchange(), then the sweep, every step (no fusion,
    # no temporal blocking: SGC_NO_FUSE-equivalent, driven through the per-call entry points)
    unfused = None
    if world == 1:
        ua, ub = A, B
        for _ in range(2):
            ua.exchange(cp)
            sgc.heat_sweep(ub, ua, a.f)
            ua, ub = ub, ua
        sgc.sync()
        sgc.profile_enable(True)
        read_profile(sgc)
        ea, eb = sgc.Event(), sgc.Event()
        nun = 20
        ea.record()
        for _ in range(nun):
            ua.exchange(cp)
            sgc.heat_sweep(ub, ua, a.f)
            ua, ub = ub, ua
        eb.record()
        sgc.sync()
        (sn, sms_), (xn, xms), _ = read_profile(sgc)
        sgc.profile_enable(False)
        t_un = ea.elapsed_ms(eb) / nun
        unfused = {"call": "sgc_exchange + sgc_heat_sweep per step (LevelData::exchange; sweep loop)", "steps": nun,
                   "ms_per_step": t_un, "value": cells_total / (t_un * 1e-3) / 1e9, "unit": "Gcell-updates/s",
                   "exchange_ms_per_sweep": xms / max(xn, 1), "sweep_ms_per_sweep": sms_ / max(sn, 1),
                   "exchange_GBs": 16.0 * cp.numCells() / (xms / max(xn, 1) * 1e-3) / 1e9 if xms > 0 else None,
                   "exchange_x_faces_from_edge_strips": ua.edgeStrips() >= 1}

    # e2e: the same job through the C ABI with HOST buffers (pinned): H2D + sweeps + D2H per step.  The host
    # side holds what an application owns — the VALID cells of every box, dense [box][k][j][i] — and moves exactly
    # those (sgc_level_upload_valid / sgc_level_download_valid): on this fully periodic level every ghost cell is
    # rewritten by the first exchange of the run, so shipping the 66^3 fabs would only add 8.8 % to both copies.
    m = a.box + 2
    vin = sgc.PinnedBuffer(A.validElems())
    vin.array[:] = pin_in.array.reshape(A.size(), m, m, m)[:, 1:-1, 1:-1, 1:-1].ravel()
    vout = sgc.PinnedBuffer(A.validElems())
    # (1) one batch at a time: upload, sweeps, download back to back (the latency of one batch)
    serial_ms = []
    for s in range(a.e2e_steps + 1 if a.e2e_steps > 0 else 0):
        barrier()
        ea, eb = sgc.Event(), sgc.Event()
        ea.record()
        A.uploadValid(vin.array)
        # B is scratch: on this fully periodic level every cell of it is written (valid cells by a sweep,
        # ghost cells by an exchange) before it is read, as with the reference's second LevelData
        res = sgc.heat_run(A, B, cp, a.f, a.sweeps)
        res.downloadValid(out=vout.array)
        eb.record()
        sgc.sync()
        if s > 0:
            serial_ms.append(ea.elapsed_ms(eb))
    serial_t = float(np.mean(serial_ms)) if serial_ms else float("nan")

    # (2) a stream of batches (the throughput the metric asks for): three level sets in HBM, batch j+1
    # uploads on one transfer stream and batch j-1 downloads on another while batch j sweeps.  Every
    # batch still crosses PCIe both ways inside the timed region; the clock runs from the first upload
    # to the last download, pipeline fill and drain included.
    pipe_t, nbatch = float("nan"), a.e2e_batches
    if a.e2e_steps > 0 and nbatch > 0:
        nset = 3
        sets = [(A, B, cp)]
        for _ in range(nset - 1):
            la, lb = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
            sets.append((la, lb, sgc.Copier().defineExchangeLD(la, 7, 0)))
        outs = [vout] + [sgc.PinnedBuffer(A.validElems()) for _ in range(nset - 1)]
        up, down = sgc.Stream(), sgc.Stream()
        ev_up = [sgc.Event() for _ in range(nset)]
        ev_run = [sgc.Event() for _ in range(nset)]
        ev_down = [sgc.Event() for _ in range(nset)]

        def submit_upload(j):
            k = j % nset
            if j >= nset:
                up.wait(ev_down[k])             # the set is free once its previous result left the device
            sets[k][0].uploadValid(vin, up)
            ev_up[k].record(up)

        def run_batches(n):
            submit_upload(0)
            for j in range(n):
                k = j % nset
                la, lb, lcp = sets[k]
                if j + 1 < n:
                    submit_upload(j + 1)
                sgc.compute_wait(ev_up[k])
                res = sgc.heat_run(la, lb, lcp, a.f, a.sweeps)      # lb is scratch (see above)
                ev_run[k].record()
                down.wait(ev_run[k])
                res.downloadValid(outs[k], down)
                ev_down[k].record(down)
            down.sync()

        run_batches(nset)                       # warm-up: staging areas, P2P sessions of the new plans
        barrier()
        ea, eb = sgc.Event(), sgc.Event()
        ea.record()
        up.wait(ea)
        run_batches(nbatch)
        eb.record(down)
        down.sync()
        pipe_t = ea.elapsed_ms(eb) / nbatch
        same = all(np.array_equal(o.array, outs[0].array) for o in outs[1:])
        if not same:
            raise SystemExit("pipelined batches disagree")
        # the platform's ceiling for this leg: the same H2D and D2H streams with NO compute between them (all ranks
        # at once): what the host side (PCIe, host DRAM, NUMA placement of the pinned buffers) can deliver
        barrier()
        ea, eb = sgc.Event(), sgc.Event()
        ea.record()
        up.wait(ea)
        down.wait(ea)
        ncopy = 8
        for j in range(ncopy):
            sets[j % nset][0].uploadValid(vin, up)
            sets[(j + 1) % nset][1].downloadValid(outs[j % nset], down)
        ec = sgc.Event()
        ec.record(up)
        down.wait(ec)
        eb.record(down)
        down.sync()
        copy_t = ea.elapsed_ms(eb) / ncopy

    copy_t = locals().get("copy_t", float("nan"))
    if dist is not None:
        import torch
        t = torch.tensor([serial_t, pipe_t, copy_t], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        serial_t, pipe_t, copy_t = float(t[0]), float(t[1]), float(t[2])
    e2e = None
    if serial_ms:
        gc = lambda ms: cells_total * a.sweeps / (ms * 1e-3) / 1e9
        piped = pipe_t == pipe_t
        e2e = {"value": gc(pipe_t if piped else serial_t), "unit": "Gcell-updates/s",
               "h2d_bytes_per_step": int(A.validElems() * 8), "d2h_bytes_per_step": int(A.validElems() * 8),
               "host_bw_ceiling_gbs": (A.validElems() * 8 / (copy_t * 1e-3) / 1e9) if copy_t == copy_t else None,
               "host_bw_ceiling": "per GPU and direction: the leg's H2D and D2H streams running concurrently with no compute, "
                                  "all ranks at once, slowest rank; ceiling_value = the metric if a step cost only that",
               "ceiling_value": gc(copy_t) if copy_t == copy_t else None, "numa_binding": numa,
               "ms_per_step": pipe_t if piped else serial_t, "pipelined": bool(piped), "batches_timed": nbatch if piped else len(serial_ms),
               "single_batch": {"value": gc(serial_t), "ms_per_step": serial_t,
                                "call": "sgc_level_upload_valid + sgc_heat_run(%d sweeps) + sgc_level_download_valid" % a.sweeps},
               "call": ("per batch: sgc_level_upload_valid_async + sgc_heat_run(%d sweeps) + sgc_level_download_valid_async "
                        "(valid cells of every box, pinned host buffers), 3 level sets, transfers of neighbouring batches "
                        "overlap the sweeps; timed from the first H2D to the last D2H" % a.sweeps) if piped else
                       "sgc_level_upload_valid + sgc_heat_run(%d sweeps) + sgc_level_download_valid, pinned host buffers" % a.sweeps}

    # Full-size parity + CPU baseline (rank 0, N = 1): the reference's own CPU run of `cpu_sweeps` sweeps of the
    # SAME level is both the checker of the device result and the timed baseline.
    cpu, parity = None, None
    restore_full_affinity()               # the CPU baseline gets every host core
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            gdom = global_domain(a, 1)
            parity, runs = parity_full_size(sgc, A, B, cp, pin_in.array, a, gdom, a.box)
            best = max(runs, key=lambda r: r["value"])
            cpu = {"value": best["value"], "unit": "Gcell-updates/s", "cores": best["cores"], "kind": best["kind"],
                   "variant": best["variant"],
                   "sample": "%d sweeps (of %d per step) of the full %dx%dx%d level in %d^3 boxes, exchange+sweep" % (
                       a.cpu_sweeps, a.sweeps, gdom[0], gdom[1], gdom[2], a.box),
                   "exchange_ms_per_sweep": best["exchange_ms"] / a.cpu_sweeps, "sweep_ms_per_sweep": best["sweep_ms"] / a.cpu_sweeps,
                   "host_cores_available": host_cores(), "all_variants": strip_fields(runs)}
        except Exception as ex:   # the baseline must never take the GPU number down with it
            cpu = {"value": None, "error": repr(ex)}
        if parity is not None and parity["ok"] is False:
            emit({"metric": "Gcell-updates/s", "value": None, "error": "device result differs from the reference", "parity": parity})
            raise SystemExit("PARITY FAILURE: %r" % (parity,))

    # the other single-GPU configs of BASELINE.json, measured in the same run (default N = 1 run only)
    configs = None
    if rank == 0 and world == 1 and a.config == "C2" and not a.no_configs:
        configs = {}
        for name, fn in (("C3", lambda: sub_config_c3(sgc, a, peak, peak_src)),
                         ("C1", lambda: run_c1(sgc, a, peak, peak_src, 5, 3, with_cpu=not a.no_cpu_baseline))):
            try:
                configs[name] = fn()
            except Exception as ex:
                configs[name] = {"error": repr(ex)}

    # ... and at 8 GPUs the 8-GPU config (collective)
    if world == 8 and a.config == "C5" and not a.no_configs:
        try:
            c4 = sub_config_c4(sgc, dist, a, world, rank, peak, peak_src)
        except Exception as ex:
            c4 = {"error": repr(ex)}
        configs = {"C4": c4}

    if rank == 0:
        name, sms, mem = sgc.device_info()
        line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": n_gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if a.config == "C4" else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(a, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "parity": parity, "parity_nranks": par_n, "unfused": unfused, "configs": configs,
                "clocks": clocks, "gpu_launches": int(launches), "device": name, "sm_count": sms,
                "plan_build_ms": plan_ms, "motion_items": cp.numMotionItem(), "remote_items": cp.numRemoteItem(),
                "rebox_factors": list(cp.reboxFactors()[1]), "pillars": list(cp.pillars()), "sum_conserved": bool(conserved)}
        emit(line)
        if par_n is not None and not par_n["ok"]:
            raise SystemExit("PARITY FAILURE at N = %d: %r" % (world, par_n))
    if dist is not None:
        dist.barrier()
        sgc.comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
