"""Build libsgc_b200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python structuredgridcalc-boxframework_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsgc_b200.so")
SOURCES = ["sgc_plan.cpp", "sgc_device.cu", "sgc_kernels.cu", "sgc_stream.cu", "sgc_stream2.cu", "sgc_comm.cu", "sgc_lb.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",              # never contract behind our back: FMAs are explicit __fma_rn calls
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-fvisibility=default",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "sgc_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    extra = os.environ.get("SGC_NVCC_EXTRA", "").split()        # A/B builds: e.g. SGC_NVCC_EXTRA=-DSGC_T2_PAIR
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(bdir, src + ".o")
        objs.append(obj)
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj] + NVCC_FLAGS + extra
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("[%s] nvcc exit code %d\n%s\n" % (src, p.returncode, out))
        ok = ok and p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"],
                   check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
