import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    import mdoracle
    mdoracle.build()
    mdoracle.lib()
    return mdoracle


@pytest.fixture(scope="session")
def md():
    import mdjl_b200
    return mdjl_b200


def force_error(F, Fref):
    """normwise per-particle error: |F_i - F_i^o|_inf / max(|F_i^o|_inf, F_rms^o)   """
    frms = np.sqrt(np.mean(Fref ** 2))
    den = np.maximum(np.max(np.abs(Fref), axis=1), max(frms, 1e-300))
    return float(np.max(np.max(np.abs(F - Fref), axis=1) / den))


def relerr(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
