"""Genuinely 16-bit test inputs (not 257 x 8-bit): shared by tests/golden/make_golden.py (which records what the reference's
own naive::compute_ssim<double, uint16_t> returns for them) and by the CPU and GPU 16-bit tests."""
import hashlib

import numpy as np

# shapes (h, w) whose reference results are recorded in tests/golden/golden.json["u16_naive"]
U16_SHAPES = [(1, 1), (3, 7), (11, 64), (64, 11), (63, 255), (65, 257), (97, 68), (40, 132), (141, 333)]
U16_FIXTURE_SHAPE = (96, 160)          # the pair stored (with its naive map) in tests/golden/u16_pair.npz


def pair16(h, w, seed, noise=3000):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = (20000 + 15000 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + rng.integers(-2000, 2001, (h, w))).clip(0, 65535)
    a = base.astype(np.uint16)
    b = (base + rng.integers(-noise, noise + 1, (h, w))).clip(0, 65535).astype(np.uint16)
    b[: h // 3, : w // 2] = a[: h // 3, : w // 2]                     # an exact-match region
    return a, b


def seed_of(h, w):
    return 7 * h + w


def digest(a, b):
    return hashlib.sha256(a.tobytes() + b.tobytes()).hexdigest()[:16]
