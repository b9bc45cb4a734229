import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracles():
    """All CPU checkers that exist on this box: the C restatement always (built on demand), the compiled
    reference when oracle/_ref/libref25519.so travelled here."""
    from oracle import pyoracle
    out = {"port": pyoracle.Oracle("port")}
    if pyoracle.available("reference"):
        out["reference"] = pyoracle.Oracle("reference")
    return out


@pytest.fixture(scope="session")
def oracle(oracles):
    """The strongest checker available: the compiled reference, else the restatement."""
    return oracles.get("reference", oracles["port"])


@pytest.fixture(scope="session")
def engine():
    if not _cuda_ok():
        pytest.skip("no CUDA device")
    import torch
    from curve25519_b200 import api
    api.init(torch.cuda.current_device())
    return api


@pytest.fixture()
def rng():
    return np.random.Generator(np.random.PCG64(0x25519))


def hx(s):
    return np.frombuffer(bytes.fromhex(s), dtype=np.uint8).copy()
