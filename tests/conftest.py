import os
import subprocess
import sys

import numpy as np
import pytest

# several slab contexts share one device in the slab parity tests and wait for each other's flags inside
# kernels: every context's streams need their own hardware queue (set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    so = os.path.join(ROOT, "particlemethod_fsi_b200", "libmphx.so")
    exe = os.path.join(ROOT, "particlemethod_fsi_b200", "Mph_Elastic_Explicit")
    if not (os.path.exists(so) and os.path.exists(exe)):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "particlemethod_fsi_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


_ensure_built()

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_PRESENT = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_2d_bar.so"))


def rel_err(a, b):
    """max-norm error relative to the max-norm of the reference array"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    d = float(np.abs(a - b).max())
    s = float(np.abs(b).max())
    return d / s if s > 0 else d


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
