"""numpy restatement of ssim_b200/csrc/synth.h; bit-identical by construction.

Used by tests and bench.py to create host-side inputs without touching the CUDA library."""
import numpy as np

DEFAULT_SEED = 0x5517
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _tri(v, p):
    m = v % (2 * p)
    return np.where(m < p, p - m, m - p)


def synth_pair(width, height, frame=0, seed=DEFAULT_SEED, y0=0):
    """Return (A, B) uint8 arrays of shape (height, width): rows y0 .. y0+height-1 of frame `frame`."""
    with np.errstate(over="ignore"):
        x = np.arange(width, dtype=np.uint64)[None, :]
        y = (np.arange(height, dtype=np.uint64) + np.uint64(y0))[:, None]
        h = _splitmix(np.uint64(seed) ^ _splitmix((np.uint64(frame) << np.uint64(40)) ^ (y << np.uint64(20)) ^ x))
    xi = x.astype(np.int64)
    yi = y.astype(np.int64)
    base = (_tri(2 * xi + yi, 256) + _tri(xi + 3 * yi, 512) // 2) // 2
    checker = (((xi >> 7) + (yi >> 7)) & 1).astype(bool)
    base = np.where(checker, base + (h & np.uint64(15)).astype(np.int64) - 8, base)
    a = np.clip(base, 0, 255)
    levels = np.array([0, 1, 2, 4, 8, 16, 32], dtype=np.int64)
    k = levels[((xi >> 8) + 3 * (yi >> 8) + frame) % 7]
    noise = ((h >> np.uint64(16)) % (2 * k + 1).astype(np.uint64)).astype(np.int64)
    b = np.clip(a + noise - k, 0, 255)
    return np.ascontiguousarray(a.astype(np.uint8)), np.ascontiguousarray(b.astype(np.uint8))


def checksum(a, b):
    """cs = cs*1099511628211 ^ a[i] ^ (b[i] << 8) over row-major pixels (u64 wrap). Slow: O(n) python ints
    avoided by a blocked Horner evaluation."""
    a = a.reshape(-1).astype(np.uint64)
    b = b.reshape(-1).astype(np.uint64)
    # xor does not distribute over multiplication, so evaluate sequentially in C-speed chunks via Python ints
    v = (a ^ (b << np.uint64(8))).tolist()
    cs = 0
    mul = 1099511628211
    mask = (1 << 64) - 1
    for t in v:
        cs = ((cs * mul) & mask) ^ t
    return cs
