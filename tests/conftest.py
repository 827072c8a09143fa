import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (`*.so` is git-ignored): build it once (nvcc cross-compiles sm_100a without a
    GPU) so that the C-ABI symbol tests and the `-m gpu` parity tests exercise the native code instead of erroring out."""
    lib = os.path.join(ROOT, "latticeqmc_b200", "liblqmc_b200.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.isfile(lib) and os.path.isfile(nvcc):
        import __graft_entry__
        __graft_entry__.build()


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.isfile(path):
        pytest.skip(f"golden fixture {name}.npz missing (run tests/golden/make_golden.py in the build container)")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture
def golden():
    return load_golden
