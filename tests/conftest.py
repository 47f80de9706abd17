import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    """The process-wide engine on cuda:0.  GPU tests FAIL (not skip) if the CUDA library is unusable."""
    from bosonic_qiskit_b200 import get_engine

    return get_engine(0)


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU should say so loudly instead of silently passing
    if not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)
