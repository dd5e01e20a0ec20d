"""pytest configuration: markers, import paths and shared fixtures.

`-m "not gpu"` : oracle vs golden vectors / the real reference, host logic, C-ABI symbol export.
`-m gpu`       : parity tests proper — every call goes through the C ABI of libsgc_b200.so.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(d, "golden.json")) as f:
        meta = json.load(f)
    arrays = np.load(os.path.join(d, "golden.npz"))
    return meta, arrays


@pytest.fixture(scope="session")
def reference():
    import pyoracle
    if not pyoracle.have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return pyoracle.Reference()


@pytest.fixture(scope="session")
def sgc():
    """The product: ctypes binding over the C ABI (fails loudly when the CUDA library is missing)."""
    from _loadpkg import load_build, load_package
    load_build().build()          # no-op when libsgc_b200.so is newer than its sources (nvcc cross-compiles without a GPU)
    return load_package()
