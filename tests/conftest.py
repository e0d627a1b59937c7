import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- fixtures shared by the GPU suites -------------------------------------------------------
@pytest.fixture(scope="module")
def D():
    import dependence_free_rl_b200 as d
    return d


@pytest.fixture(scope="module")
def ctx(D):
    c = D.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import orc as o
    o.build()
    return o
