import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ.setdefault("OMP_NUM_THREADS", "4")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """Plain-C restatement (oracle/patolette_oracle.c), compiled on demand."""
    from oracle.reflib import OracleLib
    return OracleLib()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own code: prebuilt oracle/_ref, or built here when /root/reference exists."""
    from oracle.reflib import RefLib
    try:
        return RefLib()
    except Exception as e:  # neither a prebuilt .so nor the reference tree
        pytest.skip(f"oracle/_ref unavailable: {e}")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  GPU tests call through its C ABI; missing .so is a failure, not a skip."""
    from patolette_b200 import _lib
    lib = _lib.load()
    n = lib.patolette_b200_device_count()
    assert n >= 1, f"no CUDA device visible (patolette_b200_device_count = {n})"
    return lib
