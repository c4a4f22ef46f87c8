import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def pkg():
    import maskrcnn_b200
    return maskrcnn_b200


@pytest.fixture(scope="session")
def ctx(pkg):
    """A default-configured context on cuda:0 (GPU tests only)."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = pkg.Context()
    yield c
    c.close()
