"""GPU parity tests: the CUDA path, called through the reference-facing C ABI (rmgr_ssim_compute_ssim in
librmgr-ssim.so -> ssim_cuda_compute in libssim_cuda.so), against the CPU oracle on identical inputs.

Gate (BASELINE.json north_star; reference tests/rmgr-ssim-tests.cpp:98-104):
    |global - O1| <= 2e-6   and   max |map - O1 map| <= 1e-3
where O1 is the reference's RMGR_SSIM_USE_DOUBLE build with default dispatch -- reproduced bit-for-bit on the
map by oracle.oracle_ssim(taps=TAPS_TABLE) (tests/test_oracle.py) and, when oracle/_ref is present, run live."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from conftest import GLOBAL_TOL, PIXEL_TOL
from ssim_b200 import api
from ssim_b200.synth import synth_pair

pytestmark = pytest.mark.gpu

EINSTEIN = ["einstein", "meanshift", "contrast", "impulse", "blur", "jpg"]
ERRS = []


def _report(tag, s, o, m=None, om=None):
    dg = abs(float(s) - float(o))
    line = "%-28s global %.9f  oracle %.9f  |d| %.2e" % (tag, s, o, dg)
    if m is not None:
        d = np.abs(m.astype(np.float64) - om.astype(np.float64))
        line += "   map max %.2e  p99.9 %.2e  mean %.2e" % (d.max(), np.quantile(d, 0.999), d.mean())
    ERRS.append(line)
    print(line)
    return dg


def _check_pair(tag, a, b, **kw):
    s, m = api.compute_ssim(a, b, want_map=True, **kw)
    okw = {k: v for k, v in kw.items() if k in ("width", "height", "step_a", "step_b", "stride_a", "stride_b", "a_off", "b_off")}
    o, _, om = oracle.oracle_ssim(a, b, want_map=True, taps=oracle.TAPS_TABLE, **okw)
    dg = _report(tag, s, o, m, om)
    assert dg <= GLOBAL_TOL
    assert np.abs(m - om).max() <= PIXEL_TOL
    # no-map call must give the same global value
    s2, _ = api.compute_ssim(a, b, want_map=False, **kw)
    assert s2 == s
    return s, m


def test_library_loads_and_reports_version():
    assert api.get_version() == (2, 1, 0, "2.1.0")
    assert api.cuda_lib().ssim_cuda_device_count() >= 1


@pytest.mark.parametrize("name", EINSTEIN)
def test_einstein(name, einstein, golden):
    """BASELINE.json configs[0]: the tests/images pair set, 256x256 gray, global SSIM + map"""
    s, m = _check_pair("einstein/" + name, einstein[name], einstein["einstein"])
    # the reference's own known answers (tests/rmgr-ssim-tests.cpp:354-359) at its float tolerance
    assert abs(float(s) - float(golden["einstein"][name]["golden_double_mean"])) <= GLOBAL_TOL
    if name == "einstein":
        assert s == np.float32(1.0) and (m == 1.0).all()       # identical images give exactly 1


@pytest.mark.parametrize("dims", [(1, 1), (2, 3), (7, 3), (5, 5), (11, 11), (16, 16), (255, 63), (256, 64), (257, 65), (300, 1),
                                  (1, 300), (513, 129), (64, 75), (65, 11), (130, 200), (63, 8), (59, 6), (69, 20), (70, 9), (128, 3),
                                  # width % 64 in 1..4: the right-edge patch of the second-to-last band reaches the end of its TMA box row
                                  (65, 23), (66, 23), (67, 23), (68, 97), (132, 40), (1284, 31), (1283, 12)])
def test_edge_dims(dims):
    w, h = dims
    a, b = synth_pair(w, h, 3)
    _check_pair("synthetic %dx%d" % dims, a, b)
    rng = np.random.default_rng(w * 7919 + h)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = np.clip(a.astype(np.int32) + rng.integers(-25, 26, (h, w)), 0, 255).astype(np.uint8)
    _check_pair("random %dx%d" % dims, a, b)


@pytest.mark.parametrize("dims", [(255, 63), (257, 65), (640, 80)])
def test_bbb_interleaved_crops(dims, bbb360):
    """step = 3, stride != width*step (reference tests :388-465)"""
    w, h = dims
    for ch in range(3):
        _check_pair("bbb360 %dx%d ch%d" % (w, h, ch), bbb360["jpg50"], bbb360["png"], width=w, height=h, step_a=3, step_b=3,
                    stride_a=640 * 3, stride_b=640 * 3, a_off=ch, b_off=ch)


def test_bbb_full_frames_all_channels(bbb360_full, bbb1080_green, golden):
    """The reference's bbb cases (tests/rmgr-ssim-tests.cpp:388-425) on the full frames: 640x360 PNG vs JPEG q50, every
    channel of the interleaved RGB (step = 3), and the 1080p frame (green plane); against the vectors of the unmodified
    double build recorded by make_golden.py and, when oracle/_ref is present, against that build run live."""
    png, jpg = bbb360_full["png"], bbb360_full["jpg50"]
    assert png.shape == (360, 640, 3)
    for ch in range(3):
        kw = dict(width=640, height=360, step_a=3, step_b=3, stride_a=640 * 3, stride_b=640 * 3, a_off=ch, b_off=ch)
        s, m = _check_pair("bbb360 full ch%d" % ch, jpg, png, **kw)
        g = golden["bbb360_jpg50"]["640x360_ch%d" % ch]
        assert abs(float(s) - g["ref_f64_auto"]) <= GLOBAL_TOL
        assert abs(float(m.astype(np.float64).sum()) - g["ref_f64_auto_map_sum64"]) <= GLOBAL_TOL * 640 * 360
        assert abs(float(m.min()) - g["ref_f64_auto_map_min"]) <= PIXEL_TOL
        if oracle.have_ref():
            r, rm = oracle.ref_ssim("f64", jpg, png, want_map=True, **kw)
            assert abs(float(s) - float(r)) <= GLOBAL_TOL and np.abs(m - rm).max() <= PIXEL_TOL
    # all three channels from one upload agree with the per-channel calls
    sc, mc = api.compute_channels(jpg, png, want_map=True)
    for ch in range(3):
        assert abs(float(sc[ch]) - golden["bbb360_jpg50"]["640x360_ch%d" % ch]["ref_f64_auto"]) <= GLOBAL_TOL
    png, jpg = bbb1080_green["png"], bbb1080_green["jpg50"]
    s, m = _check_pair("bbb1080 green", jpg, png)
    g = golden["bbb360_jpg50"]["1920x1080_green"]
    assert abs(float(s) - g["ref_f64_auto"]) <= GLOBAL_TOL and abs(float(m.min()) - g["ref_f64_auto_map_min"]) <= PIXEL_TOL
    if oracle.have_ref():
        r, rm = oracle.ref_ssim("f64", jpg, png, want_map=True, openmp=True)
        assert abs(float(s) - float(r)) <= GLOBAL_TOL and np.abs(m - rm).max() <= PIXEL_TOL


def test_heap_allocator_and_deprecated_overload(einstein, golden):
    """The remaining corners of the reference's test matrix (tests/rmgr-ssim-tests.cpp:468-507): Params with alloc/dealloc set
    (the "heap" variant; scratch lives in device memory here, the callbacks are accepted and never called) and the
    deprecated `float compute_ssim(const Params&)` overload, which returns the SSIM or -errno as a float (src/ssim.cpp:1108-1119)."""
    from ssim_b200._abi import make_params
    lib = api.rmgr_lib()
    a, ref = einstein["blur"], einstein["einstein"]
    want = float(golden["einstein"]["blur"]["golden_double_mean"])
    m = np.zeros((256, 256), np.float32)
    p = make_params(a, ref, 256, 256, ssim_map=m)
    assert lib.rmgr_ssim_use_default_allocator(C.byref(p)) == 0 and p.alloc and p.dealloc
    out = C.c_float()
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 0
    assert abs(out.value - want) <= GLOBAL_TOL and abs(float(m.astype(np.float64).mean()) - want) <= GLOBAL_TOL
    assert lib.rmgr_ssim_compute_ssim_openmp(C.byref(out), C.byref(p)) == 0 and abs(out.value - want) <= GLOBAL_TOL
    dep = getattr(lib, "_ZN4rmgr4ssim12compute_ssimERKNS0_6ParamsE")             # float rmgr::ssim::compute_ssim(const Params&)
    dep.restype = C.c_float
    dep.argtypes = [C.c_void_p]
    assert abs(dep(C.byref(p)) - want) <= GLOBAL_TOL
    p2 = make_params(a, ref, 256, 0)
    assert dep(C.byref(p2)) == -22.0                                              # -EINVAL as a float
    p3 = make_params(a, ref, 256, 256)
    p3.imgB.topLeft = None
    assert dep(C.byref(p3)) == -22.0


def test_synthetic_1080p_and_4k(golden):
    """BASELINE.json configs[1] (1080p, global only) and configs[2] (4K with map)"""
    a, b = synth_pair(1920, 1080, 0)
    s, _ = api.compute_ssim(a, b)
    g = golden["synthetic"]["1920x1080_f0"]
    _report("synthetic 1080p f0 (vs O1)", s, g["ref_f64_auto"])
    _report("synthetic 1080p f0 (vs O2)", s, g["ref_f64_generic"])
    assert abs(float(s) - g["ref_f64_auto"]) <= GLOBAL_TOL
    _check_pair("synthetic 1080p f0", a, b)
    a, b = synth_pair(3840, 2160, 0)
    s, m = api.compute_ssim(a, b, want_map=True)
    g = golden["synthetic"]["3840x2160_f0"]
    _report("synthetic 4K f0 (vs O1)", s, g["ref_f64_auto"])
    assert abs(float(s) - g["ref_f64_auto"]) <= GLOBAL_TOL
    if oracle.have_ref():
        r, rm = oracle.ref_ssim("f64", a, b, want_map=True, openmp=True)
        _report("synthetic 4K f0 live O1", s, r, m, rm)
        assert np.abs(m - rm).max() <= PIXEL_TOL


def test_pipelined_host_path_equals_single_shot(monkeypatch):
    """large plain-row host images go through the chunked H2D/compute/D2H pipeline; same map, same global value"""
    a, b = synth_pair(2500, 1500, 11)
    s1, m1 = api.compute_ssim(a, b, want_map=True)
    sn, _ = api.compute_ssim(a, b)
    monkeypatch.setenv("SSIM_CUDA_NO_PIPELINE", "1")
    s2, m2 = api.compute_ssim(a, b, want_map=True)
    # (centring pixels are per work item, so different row decompositions agree to rounding, not bit for bit)
    assert np.abs(m1 - m2).max() <= 3e-4 and abs(float(s1) - float(s2)) <= 2e-7 and sn == s1
    o, _, om = oracle.oracle_ssim(a, b, want_map=True, taps=oracle.TAPS_TABLE)
    assert abs(float(s1) - float(o)) <= GLOBAL_TOL and np.abs(m1 - om).max() <= PIXEL_TOL
    # row stride larger than the width (a crop of a wider image) and a map with padding
    big_a, big_b = synth_pair(3000, 1500, 11)
    buf = np.full((1500, 2600), -3.0, np.float32)
    monkeypatch.delenv("SSIM_CUDA_NO_PIPELINE")
    s3, _ = api.compute_ssim(big_a, big_b, width=2500, height=1500, stride_a=3000, stride_b=3000, ssim_map=buf, map_stride=2600)
    o = api.compute_ssim(np.ascontiguousarray(big_a[:, :2500]), np.ascontiguousarray(big_b[:, :2500]), want_map=True)
    assert abs(float(s3) - float(o[0])) <= 2e-7 and np.abs(buf[:, :2500] - o[1]).max() <= 3e-4 and (buf[:, 2500:] == -3.0).all()


def test_negative_strides_flip_and_map_layouts():
    a, b = synth_pair(97, 41, 5)
    s, m = api.compute_ssim(a, b, want_map=True)
    n = a.size
    # bottom-up images, bottom-up map with a step of 2 floats: interleaved slots must stay untouched
    buf = np.full((41, 97 * 2), -7.0, dtype=np.float32)
    sf, _ = api.compute_ssim(a, b, width=97, height=41, stride_a=-97, stride_b=-97, a_off=n - 97, b_off=n - 97,
                             ssim_map=buf, map_step=2, map_stride=-97 * 2, map_off=40 * 97 * 2)
    assert abs(float(s) - float(sf)) <= 1e-6
    # image row y was read bottom-up => output row y is image row 40-y; map written bottom-up => row 40-y of buf... = same row
    assert np.allclose(buf[:, 0::2], m, atol=2e-4) and (buf[:, 1::2] == -7.0).all()
    # column-major traversal (swap step/stride and width/height) gives the transposed map and the same global value
    st, mt = api.compute_ssim(a, b, want_map=True, width=41, height=97, step_a=97, step_b=97, stride_a=1, stride_b=1)
    assert abs(float(s) - float(st)) <= 1e-6 and np.allclose(mt, m.T, atol=2e-4)


def test_map_only_and_errors():
    a, b = synth_pair(70, 20, 1)
    s, m = api.compute_ssim(a, b, want_map=True)
    _, m2 = api.compute_ssim(a, b, want_map=True, want_ssim=False)
    assert np.array_equal(m, m2)
    lib = api.rmgr_lib()
    out = C.c_float()
    p = api.make_params(a, b, 70, 20)
    assert lib.rmgr_ssim_compute_ssim(None, C.byref(p), None) == 22                      # both outputs NULL
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), None, None) == 22                    # params NULL
    p.imgA.topLeft = None
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 22              # NULL image
    p = api.make_params(a, b, 0, 20)
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 22              # empty image (documented divergence)
    tp = api.ThreadPool()
    tp.dispatch = C.cast(C.CFUNCTYPE(C.c_int)(lambda: 0), C.c_void_p)
    tp.threadCount = 0
    p = api.make_params(a, b, 70, 20)
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), C.byref(tp)) == 22       # dispatch with 0 threads
    tp.threadCount = 4
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), C.byref(tp)) == 0        # pool accepted, not used
    assert abs(out.value - float(s)) == 0
    assert lib.rmgr_ssim_compute_ssim_openmp(C.byref(out), C.byref(p)) == 0 and out.value == float(s)


def test_absurd_size_is_enomem_not_a_crash():
    """A request whose scratch planes cannot be allocated fails with ENOMEM before any byte of the caller's images is read"""
    lib = api.cuda_lib()
    a = np.zeros(64, np.uint8)
    out = C.c_float()
    rc = lib.ssim_cuda_compute(0, 2000000000, 2000000000, a.ctypes.data, 1, 2000000000, a.ctypes.data, 1, 2000000000, None, 0, 0, C.byref(out))
    assert rc == 12, (rc, lib.ssim_cuda_last_error_string())
    s, _ = api.compute_ssim(*synth_pair(100, 50, 1))            # the library keeps working afterwards
    assert 0.0 < float(s) < 1.0


def test_flat_and_extreme_images():
    """constant images (fp32 cancellation worst case) and full-range noise"""
    for va, vb in [(0, 0), (255, 255), (200, 199), (0, 255), (17, 230)]:
        a = np.full((40, 150), va, np.uint8)
        b = np.full((40, 150), vb, np.uint8)
        s, m = api.compute_ssim(a, b, want_map=True)
        o, _, om = oracle.oracle_ssim(a, b, want_map=True, taps=oracle.TAPS_TABLE)
        dg = _report("flat %d/%d" % (va, vb), s, o, m, om)
        assert dg <= GLOBAL_TOL and np.abs(m - om).max() <= PIXEL_TOL
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (90, 333), dtype=np.uint8)
    b = rng.integers(0, 256, (90, 333), dtype=np.uint8)
    _check_pair("independent noise", a, b)


@pytest.mark.parametrize("seed", range(12))
def test_random_layouts_against_oracle(seed):
    """fuzz of the general layout path: random size, per-image step/stride (interleaved, padded, bottom-up, right-to-left,
    column-major) and map step/stride; the oracle walks the same addresses (reference ssim.h:481-499 addressing)"""
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(1, 150)), int(rng.integers(1, 90))

    def layout(buf_rng):
        kind = buf_rng.integers(0, 5)
        step = int(buf_rng.integers(1, 5))
        pad = int(buf_rng.integers(0, 9))
        if kind == 4:                                    # column-major: step spans a column
            step_b, stride_b = h * step + pad, step
            size = (w - 1) * step_b + (h - 1) * stride_b + 1
            off = 0
        else:
            row = w * step + pad
            step_b, stride_b, off = step, row, int(buf_rng.integers(0, step))
            size = h * row + step
            if kind in (1, 3):                           # bottom-up
                off += (h - 1) * row
                stride_b = -row
            if kind in (2, 3):                           # right-to-left
                off += (w - 1) * step
                step_b = -step
        return step_b, stride_b, off, size

    sa, ra, oa, na = layout(rng)
    sb, rb, ob, nb = layout(rng)
    a = rng.integers(0, 256, na, dtype=np.uint8)
    b = rng.integers(0, 256, nb, dtype=np.uint8)
    # make B a noisy copy of A at the addressed pixels so that SSIM is not ~0
    ys, xs = np.mgrid[0:h, 0:w]
    b[ob + xs * sb + ys * rb] = np.clip(a[oa + xs * sa + ys * ra].astype(int) + rng.integers(-12, 13, (h, w)), 0, 255).astype(np.uint8)
    ms = int(rng.integers(1, 4))
    mrow = w * ms + int(rng.integers(0, 5))
    flip = bool(rng.integers(0, 2))
    mbuf = np.full(h * mrow + ms, -5.0, np.float32)
    s, _ = api.compute_ssim(a, b, width=w, height=h, step_a=sa, stride_a=ra, a_off=oa, step_b=sb, stride_b=rb, b_off=ob,
                            ssim_map=mbuf, map_step=ms, map_stride=-mrow if flip else mrow, map_off=(h - 1) * mrow if flip else 0)
    o, _, om = oracle.oracle_ssim(a, b, want_map=True, width=w, height=h, step_a=sa, stride_a=ra, a_off=oa, step_b=sb, stride_b=rb, b_off=ob)
    assert abs(float(s) - float(o)) <= GLOBAL_TOL
    got = mbuf[:h * mrow].reshape(h, mrow)[:, 0:w * ms:ms]
    got = got[::-1] if flip else got
    assert np.abs(got - om).max() <= PIXEL_TOL
    touched = np.zeros(mbuf.size, bool)
    idx = ((h - 1 - ys) if flip else ys) * mrow + xs * ms
    touched[idx] = True
    assert (mbuf[~touched] == -5.0).all()                # nothing outside the addressed floats is written


def test_determinism():
    a, b = synth_pair(1000, 700, 9)
    s1, m1 = api.compute_ssim(a, b, want_map=True)
    s2, m2 = api.compute_ssim(a, b, want_map=True)
    assert s1 == s2 and np.array_equal(m1, m2)


def test_zz_error_report():
    """prints the collected error distributions (kept in gpurun_out/parity_errors.txt by the driver script)"""
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_errors.txt"), "w") as fh:
        fh.write("\n".join(ERRS) + "\n")
