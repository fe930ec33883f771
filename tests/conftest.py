import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def l1_64():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.L1(64)


@pytest.fixture(scope="session")
def l1_32():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.L1(32)


@pytest.fixture(scope="session")
def l0():
    """The compiled reference (only where oracle/_ref was built, i.e. where
    /root/reference exists or a prebuilt copy travelled with the repo)."""
    from oracle import pyoracle
    try:
        pyoracle.build()
        return pyoracle.L0()
    except Exception as e:  # pragma: no cover
        pytest.skip("reference library oracle/_ref/libref_l0.so unavailable: %s" % e)


@pytest.fixture(scope="session")
def hits_primary(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "weekend_hits_c1.npz")))


@pytest.fixture(scope="session")
def hits_bounce(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "weekend_hits_bounce.npz")))


@pytest.fixture(scope="session")
def render_c1(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "weekend_render_c1.npz")))
