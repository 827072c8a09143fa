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


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.isfile(path):
        pytest.skip(f"golden fixture {name}.npz missing (run tests/golden/make_golden.py in the build container)")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture
def golden():
    return load_golden
