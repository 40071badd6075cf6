import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rt-depth-map_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import rtdm_b200
        return rtdm_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently skip on a GPU box; without a device they are deselected by `-m "not gpu"`.
    pass


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def rt():
    import rtdm_b200
    rtdm_b200.lib()
    return rtdm_b200


@pytest.fixture(scope="session")
def gpu(rt):
    n = rt.device_count()
    assert n > 0, "no CUDA device: GPU tests must run on the B200 box (no CPU fallback exists)"
    return rt


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))
