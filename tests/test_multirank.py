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
