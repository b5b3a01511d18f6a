import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def bit_equal(a, b):
    return np.array_equal(bits(a), bits(b))


def ulp_distance(a, b):
    """Distance in units in the last place between two float32 arrays (inf/nan must match exactly)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-(2 ** 31)) - ia, ia)
    ib = np.where(ib < 0, np.int64(-(2 ** 31)) - ib, ib)
    return np.abs(ia - ib)


@pytest.fixture(scope="session")
def golden():
    return lambda name: np.load(GOLDEN / name)


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import loader
    if not loader.available("port"):
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    return loader.load("port")


@pytest.fixture(scope="session")
def ref_oracle():
    """The compiled reference; None when oracle/_ref was not built and /root/reference is absent."""
    from oracle import loader
    if not loader.available("reference") and Path("/root/reference/src").is_dir():
        subprocess.run([str(ROOT / "oracle" / "build_ref.sh")], check=True, capture_output=True)
    return loader.load("reference") if loader.available("reference") else None


@pytest.fixture(scope="session", params=["port", "reference"])
def any_oracle(request, port_oracle, ref_oracle):
    if request.param == "port":
        return port_oracle
    if ref_oracle is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return ref_oracle
