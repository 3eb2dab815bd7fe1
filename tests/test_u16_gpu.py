"""16-bit input: parity of the CUDA path (L = 65535).

The reference library hard-wires L = 255 (src/ssim.cpp:958), but its own test oracle is a template on the pixel type:
naive::compute_ssim<double, uint16_t> (tests/ssim_naive.h:230-339) IS the reference's 16-bit implementation.  The GPU path is
checked against the vectors that template produced on genuinely 16-bit inputs (tests/golden, made by make_golden.py through
oracle/_ref/libnaive.so), against the 16-bit restatement (itself pinned to the same template to 1e-12, tests/test_oracle.py),
and through the scale invariance SSIM_16(257 a, 257 b) == SSIM_8(a, b)."""
import json
import os

import numpy as np
import pytest

import u16_inputs
from conftest import GLOBAL_TOL, PIXEL_TOL
from u16_inputs import pair16 as _pair16

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", u16_inputs.U16_SHAPES)
def test_u16_matches_reference_naive_vectors(shape, golden):
    """north-star tolerances against the reference's own <double, uint16_t> template: 2e-6 global, 1e-3 per pixel"""
    from ssim_b200 import api
    h, w = shape
    a, b = _pair16(h, w, u16_inputs.seed_of(h, w))
    g = golden["u16_naive"]["%dx%d" % (w, h)]
    assert u16_inputs.digest(a, b) == g["inputs_sha256_16"]
    s, m = api.compute_u16(a, b, want_map=True)
    assert abs(float(s) - float(g["naive_double_mean"])) <= GLOBAL_TOL, (s, g["naive_double_mean"])
    assert abs(float(m.astype(np.float64).sum()) - float(g["naive_map_sum"])) <= GLOBAL_TOL * w * h
    assert abs(float(m.min()) - float(g["naive_map_min"])) <= PIXEL_TOL


def test_u16_matches_reference_naive_map_fixture(golden):
    from ssim_b200 import api
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "u16_pair.npz"))
    a, b, nm = d["a"], d["b"], d["naive_map"]
    h, w = a.shape
    s, m = api.compute_u16(a, b, want_map=True)
    assert abs(float(s) - float(golden["u16_naive"]["fixture_%dx%d" % (w, h)]["naive_double_mean"])) <= GLOBAL_TOL
    assert np.abs(m - nm).max() <= PIXEL_TOL


@pytest.mark.parametrize("shape", [(1, 1), (3, 7), (11, 64), (64, 11), (141, 333), (255, 63), (257, 65), (480, 640), (1080, 1920),
                                   (23, 65), (23, 66), (23, 67), (97, 68), (40, 132), (31, 1284)])
def test_u16_matches_oracle(shape):
    from oracle import oracle_ssim_u16
    from ssim_b200 import api
    h, w = shape
    a, b = _pair16(h, w, 7 * h + w)
    s, m = api.compute_u16(a, b, want_map=True)
    o, _, om = oracle_ssim_u16(a, b, want_map=True)
    assert abs(float(s) - float(o)) <= GLOBAL_TOL, (s, o)
    assert np.abs(m - om).max() <= PIXEL_TOL
    s2, _ = api.compute_u16(a, b, want_map=False)
    assert abs(float(s2) - float(s)) <= 2e-7


def test_u16_scale_invariance_pins_the_8_bit_path():
    from ssim_b200 import api
    tests_dir = __import__("os").path.dirname(__file__)
    g = np.load(tests_dir + "/golden/einstein.npz")
    ref = g["einstein"]
    import json
    golden = json.load(open(tests_dir + "/golden/golden.json"))["einstein"]
    for name in ("blur", "contrast", "impulse", "jpg", "meanshift"):
        img = g[name]
        s16, m16 = api.compute_u16(ref.astype(np.uint16) * 257, img.astype(np.uint16) * 257, want_map=True)
        s8, m8 = api.compute_ssim(ref, img, want_map=True)
        assert abs(float(s16) - float(s8)) <= 5e-7 and np.abs(m16 - m8).max() <= 2e-4      # a few float ulps: C1, C2 and the partition of the work round differently
        assert abs(float(s16) - float(golden[name]["golden_double_mean"])) <= GLOBAL_TOL   # the reference's own known answers (tests/rmgr-ssim-tests.cpp:354-359)


def test_u16_identical_flat_and_extremes():
    from ssim_b200 import api
    a, _ = _pair16(97, 203, 5)
    s, m = api.compute_u16(a, a.copy(), want_map=True)
    assert s == np.float32(1.0) and (m == 1.0).all()
    lo = np.zeros((40, 90), np.uint16); hi = np.full((40, 90), 65535, np.uint16)
    from oracle import oracle_ssim_u16
    for x, y in ((lo, hi), (hi, lo), (hi, hi), (lo, lo)):
        s, m = api.compute_u16(x, y, want_map=True)
        o, _, om = oracle_ssim_u16(x, y, want_map=True)
        assert abs(float(s) - float(o)) <= GLOBAL_TOL and np.abs(m - om).max() <= PIXEL_TOL


def test_u16_strided_layouts_and_device_api():
    import torch
    from oracle import oracle_ssim_u16
    from ssim_b200 import api
    h, w = 120, 200
    a, b = _pair16(h, w, 11)
    o, _, om = oracle_ssim_u16(a, b, want_map=True)
    # interleaved 3-channel, channel 1; and bottom-up rows
    ia = np.zeros((h, w, 3), np.uint16); ib = np.zeros((h, w, 3), np.uint16)
    ia[..., 1] = a; ib[..., 1] = b
    s, m = api.compute_u16(ia, ib, want_map=True, width=w, height=h, step_a=3, step_b=3, stride_a=3 * w, stride_b=3 * w, a_off=1, b_off=1)
    assert abs(float(s) - float(o)) <= GLOBAL_TOL and np.abs(m - om).max() <= PIXEL_TOL
    fa = np.ascontiguousarray(a[::-1]); fb = np.ascontiguousarray(b[::-1])
    s, _ = api.compute_u16(fa, fb, width=w, height=h, stride_a=-w, stride_b=-w, a_off=(h - 1) * w, b_off=(h - 1) * w)
    assert abs(float(s) - float(o)) <= GLOBAL_TOL
    # device-resident batch of 3 frames through ssim_cuda_compute_device_u16
    pitch = (2 * w + 15) // 16 * 16
    frames = 3
    da = torch.zeros((frames, h, pitch // 2), dtype=torch.int16, device="cuda"); db = torch.zeros_like(da)
    for f in range(frames):
        da[f, :, :w] = torch.from_numpy(np.roll(a, 5 * f, axis=1).view(np.int16)).cuda()
        db[f, :, :w] = torch.from_numpy(np.roll(b, 5 * f, axis=1).view(np.int16)).cuda()
    dm = torch.empty((frames, h, w), dtype=torch.float32, device="cuda")
    ds = torch.empty(frames, dtype=torch.float32, device="cuda")
    api.compute_device_u16(0, None, w, h, 0, h, frames, da.data_ptr(), pitch, pitch * h, db.data_ptr(), pitch, pitch * h,
                           dm.data_ptr(), w, w * h, None, ds.data_ptr())
    torch.cuda.synchronize()
    for f in range(frames):
        of, _, omf = oracle_ssim_u16(np.roll(a, 5 * f, axis=1).copy(), np.roll(b, 5 * f, axis=1).copy(), want_map=True)
        assert abs(float(ds[f]) - float(of)) <= GLOBAL_TOL and np.abs(dm[f].cpu().numpy() - omf).max() <= PIXEL_TOL


def test_u16_pipelined_host_path_equals_single_shot(monkeypatch):
    """Large plain-row host images take the chunked H2D / kernel / D2H pipeline; results must match the single-shot path
    (up to the per-item centring, which follows the chunk boundaries) and the oracle."""
    from oracle import oracle_ssim_u16
    from ssim_b200 import api
    a, b = _pair16(1200, 1600, 3)
    s1, m1 = api.compute_u16(a, b, want_map=True)                      # pipelined (>= 2^20 pixels)
    monkeypatch.setenv("SSIM_CUDA_NO_PIPELINE", "1")
    s2, m2 = api.compute_u16(a, b, want_map=True)                      # single shot
    monkeypatch.delenv("SSIM_CUDA_NO_PIPELINE")
    o, _, om = oracle_ssim_u16(a, b, want_map=True)
    assert np.abs(m1 - m2).max() <= 3e-4 and abs(float(s1) - float(s2)) <= 2e-7
    assert abs(float(s1) - float(o)) <= GLOBAL_TOL and np.abs(m1 - om).max() <= PIXEL_TOL
    sn, _ = api.compute_u16(a, b, want_map=False)
    assert abs(float(sn) - float(s1)) <= 2e-7
