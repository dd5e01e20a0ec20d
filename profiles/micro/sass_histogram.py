"""Static SASS opcode histograms of the kernels on the hot path (cuobjdump -sass of the built library).
    python profiles/micro/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "structuredgridcalc-boxframework_b200", "libsgc_b200.so")
WANT = (("k_heat_t2<64, NS 4, 16 warps, fma, periodic>", "k_heat_t2ILi64ELi4ELi16ELb1ELb0"),
        ("k_heat_t2<64, ..., physical boundaries>", "k_heat_t2ILi64ELi4ELi16ELb1ELb1"),
        ("k_heat_stream<64,64, PS 1, NS 6, 16 warps, double> (plain one-step sweep)", "k_heat_streamILi64ELi64ELi1ELi1ELi6ELi16ELi0ELb0ELi0ELb0Ed"),
        ("k_heat_stream<64,64, PS 2, NS 6, 16 warps, float>", "k_heat_streamILi64ELi64ELi1ELi2ELi6ELi16ELi0ELb0ELi0ELb0Ef"),
        ("k_heat_stream<64,64, ..., fused ghost fill (pull y/z, push x)>", "k_heat_streamILi64ELi64ELi1ELi1ELi6ELi16ELi0ELb1ELi0ELb0Ed"),
        ("k_copy_items<u64> (batched region copy: exchange, pack / unpack, BaseFab::copy, BC)", "k_copy_itemsImEE"),
        ("k_sweep7<double, LapOp> (laplacian driver, C1)", "k_sweep7IdNS_5LapOpIdEELi1"))

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, hist = None, collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        hist[cur][m.group(1).split(".")[0]] += 1
print("SASS opcode histograms: static instruction counts, cuobjdump -sass of libsgc_b200.so built from the committed sources")
print("(UBLKCP = cp.async.bulk, the 1-D TMA copy; SYNCS = mbarrier ops; LDGSTS = cp.async; no UTMALDG: no tensor maps needed;")
print(" no HMMA/UTCMMA: bandwidth-bound stencil, tensor cores unused by design)\n")
for title, key in WANT:
    for name, h in hist.items():
        if key in name:
            tot = sum(h.values())
            print("%s\n  %s" % (title, name))
            print("  %d instructions: %s" % (tot, ", ".join("%s %d" % (k, v) for k, v in h.most_common(24))))
            print("  bulk copy / mbarrier: UBLKCP %d, SYNCS %d, LDGSTS %d | FP64: DADD %d DMUL %d DFMA %d | FP32: FADD %d FMUL %d FFMA %d | "
                  "LDS %d STS %d LDG %d STG %d | IMAD.* %d MOV %d\n" % (h["UBLKCP"], h["SYNCS"], h["LDGSTS"], h["DADD"], h["DMUL"], h["DFMA"],
                                                                      h["FADD"], h["FMUL"], h["FFMA"], h["LDS"], h["STS"], h["LDG"], h["STG"],
                                                                      h["IMAD"], h["MOV"]))
            break
