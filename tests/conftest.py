import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_sessionstart(session):
    """Build what the tests load if it is not there yet (fresh clone): the CUDA library (nvcc
    cross-compiles sm_100a without a GPU), the oracle and the C++ adapter."""
    so = os.path.join(ROOT, "tweakseq_b200", "libtsqb200.so")
    oracle_so = os.path.join(ROOT, "oracle", "libtsq_oracle.so")
    if not (os.path.exists(so) and os.path.exists(oracle_so)):
        import __graft_entry__
        __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run under gpurun); everything else runs on CPU")


def _has_gpu():
    try:
        import tweakseq_b200 as t
        return t.load_library().tsq_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a B200 must fail loudly rather than silently skip the parity tests
    if config.getoption("-m") and "gpu" in config.getoption("-m") and "not gpu" not in config.getoption("-m"):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no sm_100 device here")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
