"""Shared test plumbing.

CPU tests (`-m "not gpu"`) cover the oracle against the committed golden vectors, the host
scheduler mirror and the C-ABI surface.  GPU tests (`-m gpu`) are the parity tests proper:
they call libqrb200.so through its C ABI on device memory and compare with the oracle.
Nothing here reads /root/reference at run time (the GPU box does not have it).
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN, "ref_small.npz"))


@pytest.fixture(scope="session")
def po():
    import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def qr():
    import gpu_tiled_qr_decomposition_b200 as q
    q.build()
    return q


@pytest.fixture(scope="session")
def cuda(qr):
    """torch on cuda:0 (device memory + streams only) and the loaded library."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    torch.cuda.set_device(0)
    qr.lib()
    return torch
