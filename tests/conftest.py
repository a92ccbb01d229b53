import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
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
def golden_prims():
    import json
    with open(os.path.join(GOLDEN, "prims.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLDEN, "prims.npz"))


@pytest.fixture(scope="session")
def golden_nets():
    return np.load(os.path.join(GOLDEN, "nets.npz"))


@pytest.fixture(scope="session")
def golden_cells():
    return np.load(os.path.join(GOLDEN, "cells.npz"))


@pytest.fixture(scope="session")
def lib():
    from nas_3d_unet_b200 import build, _lib
    build.build_library()
    return _lib.load()
