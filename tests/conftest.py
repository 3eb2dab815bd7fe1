import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# Tolerances of the north star (BASELINE.json) == the reference's float-build test tolerances
# (reference tests/rmgr-ssim-tests.cpp:98-104): measured against the RMGR_SSIM_USE_DOUBLE build.
GLOBAL_TOL = 2e-6
PIXEL_TOL = 1e-3


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def einstein():
    return dict(np.load(os.path.join(GOLDEN_DIR, "einstein.npz")))


@pytest.fixture(scope="session")
def bbb360():
    return dict(np.load(os.path.join(GOLDEN_DIR, "bbb360_top80.npz")))
