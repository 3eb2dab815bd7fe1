"""Pins the CPU oracle (oracle/ssim_oracle.c) -- CPU only, no GPU needed.

 1. against the reference's own decoder-independent known answers (tests/rmgr-ssim-tests.cpp:354-359),
 2. against golden vectors produced by the unmodified reference build and committed in tests/golden/,
 3. live against oracle/_ref (the reference compiled in place) when those libraries are present."""
import numpy as np
import pytest

import oracle
from ssim_b200.synth import checksum, synth_pair

EINSTEIN = ["einstein", "meanshift", "contrast", "impulse", "blur", "jpg"]
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference once)")


def test_taps_are_normalised_and_symmetric():
    k = oracle.oracle_taps(oracle.TAPS_RUNTIME)
    assert abs(k.sum() - 1.0) < 1e-15
    assert np.array_equal(k, k.T) and np.array_equal(k, k[::-1, ::-1])
    # separable up to double rounding: k == outer(g, g)
    g = k.sum(axis=0)
    assert np.abs(np.outer(g, g) - k).max() < 1e-16
    t = oracle.oracle_taps(oracle.TAPS_TABLE)
    # the float-pipeline table is biased: sum = 1 + 1.02e-8
    assert 0.9e-8 < t.sum() - 1.0 < 1.2e-8
    assert np.array_equal(t, t.T)


@pytest.mark.parametrize("name", EINSTEIN)
def test_einstein_known_answers(name, einstein, golden):
    """reference tests/rmgr-ssim-tests.cpp:354-359, REF_TOLERANCE 1e-13 (:64-73)"""
    a, ref = einstein[name], einstein["einstein"]
    _, total, _ = oracle.oracle_ssim(a, ref, taps=oracle.TAPS_RUNTIME)
    known = float(golden["einstein"][name]["golden_double_mean"])
    assert abs(total / (256 * 256) - known) < 1e-13


@pytest.mark.parametrize("name", EINSTEIN)
def test_einstein_matches_reference_build_vectors(name, einstein, golden):
    a, ref = einstein[name], einstein["einstein"]
    g = golden["einstein"][name]
    s1, _, m1 = oracle.oracle_ssim(a, ref, want_map=True, taps=oracle.TAPS_TABLE)
    s2, _, _ = oracle.oracle_ssim(a, ref, taps=oracle.TAPS_RUNTIME)
    assert abs(float(s1) - g["ref_f64_auto"]) <= 6e-8        # one float ulp
    assert abs(float(s2) - g["ref_f64_generic"]) <= 6e-8
    assert abs(m1.astype(np.float64).sum() - g["ref_f64_auto_map_sum64"]) < 1e-6


def test_einstein_map_fixture(einstein):
    import os
    from conftest import GOLDEN_DIR
    want = np.load(os.path.join(GOLDEN_DIR, "einstein_blur_map_f64auto.npz"))["map"]
    _, _, got = oracle.oracle_ssim(einstein["blur"], einstein["einstein"], want_map=True, taps=oracle.TAPS_TABLE)
    assert np.abs(got - want).max() <= 1.2e-7


@pytest.mark.parametrize("dims", [(255, 63), (257, 65), (640, 80)])
@pytest.mark.parametrize("ch", [0, 1, 2])
def test_bbb_interleaved_crops(dims, ch, bbb360, golden):
    """step = 3, stride != width*step, one-tile-minus/plus-one shapes (reference tests :428-465)"""
    w, h = dims
    g = golden["bbb360_jpg50"]["%dx%d_ch%d" % (w, h, ch)]
    s, _, m = oracle.oracle_ssim(bbb360["jpg50"], bbb360["png"], want_map=True, taps=oracle.TAPS_TABLE,
                                 step_a=3, step_b=3, stride_a=640 * 3, stride_b=640 * 3, width=w, height=h, a_off=ch, b_off=ch)
    assert abs(float(s) - g["ref_f64_auto"]) <= 6e-8
    assert abs(m.astype(np.float64).sum() - g["ref_f64_auto_map_sum64"]) < 1e-6


def test_synthetic_recipe_checksums(golden):
    for key in ["1920x1080_f0", "1920x1080_f1"]:
        w, h = map(int, key.split("_")[0].split("x"))
        f = int(key.split("_f")[1])
        a, b = synth_pair(w, h, f)
        assert "%016x" % checksum(a, b) == golden["synthetic"][key]["checksum"]
    # strips of a frame are slices of the frame
    a, b = synth_pair(300, 40, 2)
    a2, b2 = synth_pair(300, 15, 2, y0=25)
    assert np.array_equal(a[25:], a2) and np.array_equal(b[25:], b2)


@pytest.mark.parametrize("key", ["1x1_f3", "2x3_f3", "7x3_f3", "5x5_f3", "11x11_f3", "16x16_f3", "255x63_f3", "256x64_f3",
                                 "257x65_f3", "300x1_f3", "1x300_f3", "513x129_f3", "64x75_f3", "65x11_f3", "130x200_f3"])
def test_edge_dims_vectors(key, golden):
    w, h = map(int, key.split("_")[0].split("x"))
    a, b = synth_pair(w, h, 3)
    g = golden["synthetic"][key]
    s1, _, m = oracle.oracle_ssim(a, b, want_map=True, taps=oracle.TAPS_TABLE)
    s2, _, _ = oracle.oracle_ssim(a, b, taps=oracle.TAPS_RUNTIME)
    assert abs(float(s1) - g["ref_f64_auto"]) <= 6e-8
    assert abs(float(s2) - g["ref_f64_generic"]) <= 6e-8
    assert abs(m.astype(np.float64).sum() - g["ref_f64_auto_map_sum64"]) < 1e-7 * max(1, w * h)


def test_synthetic_1080p_vector(golden):
    a, b = synth_pair(1920, 1080, 0)
    s1, _, _ = oracle.oracle_ssim(a, b, taps=oracle.TAPS_TABLE)
    assert abs(float(s1) - golden["synthetic"]["1920x1080_f0"]["ref_f64_auto"]) <= 6e-8


def test_argument_errors():
    import ctypes as C
    lib = oracle.oracle_lib()
    a = np.zeros((4, 4), np.uint8)
    s = C.c_float()
    assert lib.ssim_oracle_compute(4, 4, a.ctypes.data, 1, 4, a.ctypes.data, 1, 4, None, 0, 0, 0, None, None) == 22
    assert lib.ssim_oracle_compute(4, 4, None, 1, 4, a.ctypes.data, 1, 4, None, 0, 0, 0, C.byref(s), None) == 22
    assert lib.ssim_oracle_compute(0, 4, a.ctypes.data, 1, 4, a.ctypes.data, 1, 4, None, 0, 0, 0, C.byref(s), None) == 22


def test_negative_strides_and_identity():
    """flipping both images leaves the global SSIM unchanged; identical images give exactly 1"""
    a, b = synth_pair(97, 41, 5)
    s, _, m = oracle.oracle_ssim(a, b, want_map=True)
    n = a.size
    sf, _, mf = oracle.oracle_ssim(a, b, want_map=True, stride_a=-97, stride_b=-97, a_off=n - 97, b_off=n - 97, width=97, height=41)
    assert abs(float(s) - float(sf)) < 1e-7 and np.abs(m[::-1] - mf).max() < 1e-6
    s1, _, m1 = oracle.oracle_ssim(a, a, want_map=True)
    assert s1 == np.float32(1.0) and (m1 == 1.0).all()


# ------------------------------------------------------------------ live against the reference build
@needs_ref
@pytest.mark.parametrize("dims", [(1, 1), (3, 2), (33, 9), (256, 64), (300, 70), (517, 131)])
def test_live_against_reference_random(dims):
    w, h = dims
    rng = np.random.default_rng(w * 1000 + h)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = np.clip(a.astype(np.int32) + rng.integers(-20, 21, (h, w)), 0, 255).astype(np.uint8)
    s1, _, m1 = oracle.oracle_ssim(a, b, want_map=True, taps=oracle.TAPS_TABLE)
    r1, rm1 = oracle.ref_ssim("f64", a, b, want_map=True)
    assert abs(float(s1) - float(r1)) <= 6e-8 and np.abs(m1 - rm1).max() <= 1.2e-7
    s2, _, m2 = oracle.oracle_ssim(a, b, want_map=True, taps=oracle.TAPS_RUNTIME)
    r2, rm2 = oracle.ref_ssim("f64", a, b, want_map=True, impl=oracle.IMPL_GENERIC)
    assert abs(float(s2) - float(r2)) <= 6e-8 and np.abs(m2 - rm2).max() <= 1.2e-7


@needs_ref
def test_reference_builds_are_what_they_claim():
    assert oracle.ref_lib("f32").ref_uses_double() == 0
    assert oracle.ref_lib("f64").ref_uses_double() == 1
    # AUTO on x86 picks FMA when available: mask has generic|auto at least
    assert oracle.ref_lib("f32").ref_select_impl(0) & 0x3 == 0x3


def test_u16_oracle_scale_invariance():
    """Scale invariance of the 16-bit restatement (L = 65535): SSIM_16(257 a, 257 b) == SSIM_8(a, b), because C1 and C2 scale
    with L^2 and 65535 = 257 * 255.  (The pin to the reference's own 16-bit template is further down.)"""
    from oracle import oracle_ssim, oracle_ssim_u16
    rng = np.random.default_rng(3)
    for h, w in ((1, 1), (7, 19), (97, 131)):
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-25, 26, a.shape), 0, 255).astype(np.uint8)
        for taps in (0, 1):
            s8, d8, m8 = oracle_ssim(a, b, want_map=True, taps=taps)
            s16, d16, m16 = oracle_ssim_u16(a.astype(np.uint16) * 257, b.astype(np.uint16) * 257, want_map=True, taps=taps)
            assert abs(d8 - d16) <= 1e-9 * max(1.0, abs(d8)) and np.abs(m8 - m16).max() <= 1e-6 and abs(float(s8) - float(s16)) <= 1e-7
    # genuinely 16-bit content: symmetric, 1.0 on identical images, below 1 otherwise
    a = rng.integers(0, 65536, (33, 47), dtype=np.uint16)
    b = np.clip(a.astype(int) + rng.integers(-3000, 3001, a.shape), 0, 65535).astype(np.uint16)
    sab, _, _ = oracle_ssim_u16(a, b)
    sba, _, _ = oracle_ssim_u16(b, a)
    saa, _, maa = oracle_ssim_u16(a, a.copy(), want_map=True)
    assert sab == sba and saa == np.float32(1.0) and (maa == 1.0).all() and sab < 1.0


# ---- 16-bit pinned to the reference's own naive::compute_ssim<double, uint16_t> (tests/ssim_naive.h:230-339)
import u16_inputs  # noqa: E402


@pytest.mark.parametrize("shape", u16_inputs.U16_SHAPES)
def test_u16_oracle_matches_reference_naive_vectors(shape, golden):
    """Committed vectors: what oracle/_ref/libnaive.so (the reference's template, T = uint16_t => L = 65535) returned for
    genuinely 16-bit inputs; the restatement with true-math taps must agree to 1e-12 on the double mean."""
    h, w = shape
    a, b = u16_inputs.pair16(h, w, u16_inputs.seed_of(h, w))
    g = golden["u16_naive"]["%dx%d" % (w, h)]
    assert u16_inputs.digest(a, b) == g["inputs_sha256_16"], "numpy's generator no longer reproduces the recorded inputs"
    _, total, m = oracle.oracle_ssim_u16(a, b, want_map=True, taps=oracle.TAPS_RUNTIME)
    assert abs(total / (w * h) - float(g["naive_double_mean"])) <= 1e-12
    assert abs(float(m.min()) - float(g["naive_map_min"])) <= 1e-6            # the oracle's map is float32
    # the gating oracle mode (float-pipeline window, as every shipped reference build applies it) stays within the
    # window's normalisation bias of the true-math one
    s_tab, _, _ = oracle.oracle_ssim_u16(a, b, taps=oracle.TAPS_TABLE)
    assert abs(float(s_tab) - float(g["naive_double_mean"])) <= 2e-6


def test_u16_oracle_matches_reference_naive_map_fixture(golden):
    d = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "u16_pair.npz"))
    a, b, nm = d["a"], d["b"], d["naive_map"]
    assert a.dtype == np.uint16 and int(a.max()) > 255                          # genuinely 16-bit content
    h, w = a.shape
    _, total, m = oracle.oracle_ssim_u16(a, b, want_map=True, taps=oracle.TAPS_RUNTIME)
    assert abs(total / (w * h) - float(golden["u16_naive"]["fixture_%dx%d" % (w, h)]["naive_double_mean"])) <= 1e-12
    assert np.abs(m - nm).max() <= 2e-7                                         # both maps rounded to float32


@pytest.mark.skipif(not oracle.have_naive(), reason="oracle/_ref/libnaive.so not built (needs /root/reference once)")
def test_u16_and_u8_oracle_live_against_reference_naive():
    rng = np.random.default_rng(11)
    for h, w in ((5, 9), (64, 64), (70, 131)):
        a = rng.integers(0, 65536, (h, w), dtype=np.uint16)
        b = np.clip(a.astype(int) + rng.integers(-5000, 5001, a.shape), 0, 65535).astype(np.uint16)
        mean, nm = oracle.naive_ssim(a, b, want_map=True)
        _, total, m = oracle.oracle_ssim_u16(a, b, want_map=True, taps=oracle.TAPS_RUNTIME)
        assert abs(total / (w * h) - mean) <= 1e-12 and np.abs(m - nm.astype(np.float32)).max() <= 2e-7
        a8, b8 = (a >> 8).astype(np.uint8), (b >> 8).astype(np.uint8)
        mean8, _ = oracle.naive_ssim(a8, b8)
        _, total8, _ = oracle.oracle_ssim(a8, b8, taps=oracle.TAPS_RUNTIME)
        assert abs(total8 / (w * h) - mean8) <= 1e-12
