import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.set_math(pyoracle.MATH_CONTRACT)
    return pyoracle


def pytest_collection_modifyitems(config, items):
    """On a host without CUDA (or without the built library) the gpu-marked tests are skipped with the reason,
    instead of failing in FitnessEngine ("no CPU fallback") -- `pytest tests` is green on a CPU box."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_gpu = False
    have_lib = os.path.exists(os.path.join(ROOT, "multitreegp_b200", "libmtgp_b200.so"))
    if have_gpu and have_lib:
        return
    why = "needs a CUDA device" if not have_gpu else "libmtgp_b200.so is not built (python __graft_entry__.py)"
    skip = pytest.mark.skip(reason=why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
