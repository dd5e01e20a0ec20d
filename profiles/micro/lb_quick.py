import json, os, sys
sys.path.insert(0, os.getcwd())
from _loadpkg import load_package
sgc = load_package(); sgc.init(0)
for (nx, ny, nz, bs, nstep) in ((256, 128, 128, 32, 60), (512, 256, 256, 64, 20)):
    dbl = sgc.DisjointBoxLayout(sgc.Box((0, 0, 0), (nx - 1, ny - 1, nz - 1)), (bs, bs, bs))
    lb = sgc.LBLevel(dbl); lb.advance(5, fuse=True); sgc.sync()
    e0, e1 = sgc.Event(), sgc.Event(); e0.record(); lb.advance(nstep, fuse=True); e1.record()
    ms = e0.elapsed_ms(e1)
    print("  %dx%dx%d box %d fused: %.3f ms/step  %.2f Gcell/s" % (nx, ny, nz, bs, ms / nstep, nx * ny * nz * nstep / (ms * 1e-3) / 1e9), flush=True)
    del lb
