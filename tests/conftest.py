import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ref_mc():
    """the reference's own compiled Cython marching cubes (oracle/_ref), or None when it was not built"""
    p = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(p, "meshudf")):
        return None
    sys.path.insert(0, p)
    try:
        from meshudf import _marching_cubes_lewiner_cy as cy
        from meshudf._marching_cubes_lewiner import _get_mc_luts
    except Exception:
        return None
    luts = _get_mc_luts()
    return lambda udf, g: cy.marching_cubes_udf(udf, g, luts, 1, 0, None)[:2]
