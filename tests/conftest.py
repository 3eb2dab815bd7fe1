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


sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing them (there is no CPU fallback
    to run them on)."""
    try:
        from ssim_b200 import api
        have_gpu = api.cuda_lib().ssim_cuda_device_count() > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device: ssim_b200 has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def einstein():
    return dict(np.load(os.path.join(GOLDEN_DIR, "einstein.npz")))


@pytest.fixture(scope="session")
def bbb360_full():
    """big_buck_bunny_360_07806 PNG and JPEG q50 (decoded by ssim_b200/csrc/jpeg_reader.h: the pixels the reference's loader makes), full 640x360 RGB frames"""
    return dict(np.load(os.path.join(GOLDEN_DIR, "bbb360.npz")))


@pytest.fixture(scope="session")
def bbb360(bbb360_full):
    """the top 80 rows of the same frames (enough for the reference's 255x63 / 257x65 crops)"""
    return {k: np.ascontiguousarray(v[:80]) for k, v in bbb360_full.items()}


@pytest.fixture(scope="session")
def bbb1080_green():
    return dict(np.load(os.path.join(GOLDEN_DIR, "bbb1080_green.npz")))
