import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 GPU (run with -m gpu on the B200 box)")


def golden(name):
    with np.load(os.path.join(GOLDEN, "golden_%s.npz" % name)) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def lib():
    """libgpemu.so, built on demand (nvcc cross-compiles without a GPU)."""
    from gp_emulator_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-j8"], cwd=ROOT)
    return _lib.load()
