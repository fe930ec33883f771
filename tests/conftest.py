import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# Property tests (hypothesis) draw the SAME examples on every run by default, so that the suite is repeatable (a red
# test is a bug, not luck); RT_HYPOTHESIS_EXPLORE=1 switches to fresh random examples - the development loop that found
# e.g. the box whose middle rects belong to a group (tests/test_random_scenes.py).
try:
    from hypothesis import settings as _hyp_settings
    _hyp_settings.register_profile("repeatable", derandomize=True)
    _hyp_settings.register_profile("explore", derandomize=False)
    _hyp_settings.load_profile("explore" if os.environ.get("RT_HYPOTHESIS_EXPLORE") else "repeatable")
except ImportError:  # hypothesis is optional for the rest of the suite
    pass


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
