"""world_size > 1: the host-side multi-rank logic on CPU (gloo, 2 and 4 ranks) and — where >= 2
GPUs are visible — the NCCL halo path of the product, bit-exact against the oracle."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(mode, nproc, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("nproc", [2, 4])
def test_multirank_host_logic_gloo(nproc):
    out = _launch("cpu", nproc)
    assert out.count("cpu multi-rank host logic ok") == nproc


@pytest.mark.gpu
def test_multirank_nccl_halo_parity():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    out = _launch("gpu", 2)
    assert out.count("gpu multi-rank parity ok") == 2
    if n >= 4:          # several peers per rank: per-peer message order and unpack-as-arrived
        out = _launch("gpu", 4)
        assert out.count("gpu multi-rank parity ok") == 4


@pytest.mark.gpu
def test_cpp_facade_driver_two_ranks(oracle, tmp_path):
    """apps/levelheat (written against the C++ facade, DisjointBoxLayout::initMPI + NCCL id through
    $SGC_COMM_FILE) as two processes on two GPUs; each rank's hash of its boxes against the oracle."""
    import json
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    exe = os.path.join(ROOT, "apps", "_build", "levelheat")
    if not os.path.exists(exe):
        pytest.skip("apps/_build/levelheat not built")
    n, bs, nstep, world = 64, 32, 20, 2
    procs = []
    for r in range(world):
        env = dict(os.environ, SGC_RANK=str(r), SGC_NPROC=str(world), SGC_DEVICE=str(r),
                   SGC_COMM_FILE=str(tmp_path / "ncclid"))
        procs.append(subprocess.Popen([exe, str(n), str(bs), str(nstep)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    dlo, dhi, mb = (0, 0, 0), (n - 1, n - 1, n * world - 1), (bs,) * 3
    boxes, owner, _ = oracle.layout(dlo, dhi, mb, world)
    u = oracle.fill_hash(boxes, 1, 0.0)
    oracle.level_heat(dlo, dhi, mb, 1, 7, 0, 0.125, nstep, u, 1)
    m = bs + 2
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out + err
        line = json.loads(out.strip().splitlines()[-1])
        assert line["rank"] == r and line["nproc"] == world
        chunks = [u[b * m ** 3:(b + 1) * m ** 3].reshape(m, m, m)[1:-1, 1:-1, 1:-1].ravel()
                  for b in range(len(boxes)) if owner[b] == r]
        assert line["local_fnv1a64"] == "%016x" % oracle.fnv1a64(np.concatenate(chunks)), (r, line)
