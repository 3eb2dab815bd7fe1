"""The reference's JPEG-based suites (tests/rmgr-ssim-tests.cpp:364-465: bbb360, bbb1080, bbb255, bbb257 -- a PNG frame
against eleven JPEG encodings of it, every channel) with the reference's OWN known answers.

Those 132 constants are SSIMs of the pixels the reference's image loader (stb_image) makes of the JPEG files, so they bind
only a decoder with identical pixels: libjpeg's differ in ~1 % of the samples by one level, which moves the means by up to
8e-6 -- four times the reference's tolerance.  The front end's own reader (ssim_b200/csrc/jpeg_reader.h) reproduces them:

  CPU  the reader + the reference's naive template (oracle/_ref/libnaive.so) or our C restatement -> the constants to 1e-13
       (REF_TOLERANCE of the reference, :72), i.e. the decoded pixels are identical; this also pins the oracle on real content
  GPU  the reference's test itself: interleaved RGB, step 3, full-frame stride, width/height cropped, every channel and
       quality, result against the constant at the reference's float tolerance 2e-6 (:102), map against the oracle at 1e-3

The 360p files travel as bytes in tests/golden/bbb360_jpeg_files.npz (0.7 MB); the 1080p files (3.8 MB) are read from
/root/reference where it exists (CPU tests only; skipped elsewhere)."""
import io
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import GLOBAL_TOL, GOLDEN_DIR, PIXEL_TOL
from ssim_b200 import api

REF_IMAGES = "/root/reference/tests/images"
REF_TOLERANCE = 1e-13
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ssim_b200", "bin", "rmgr-ssim")
# suite -> (width, height) the reference passes as maxWidth / maxHeight (tests/rmgr-ssim-tests.cpp:405, 445, 465)
CROPS = {"bbb360": (640, 360), "bbb255": (255, 63), "bbb257": (257, 65)}


@pytest.fixture(scope="module")
def jpeg_files():
    return dict(np.load(os.path.join(GOLDEN_DIR, "bbb360_jpeg_files.npz")))


@pytest.fixture(scope="module")
def decoded(jpeg_files):
    return {int(k[1:]): api.decode_jpeg(v) for k, v in jpeg_files.items()}


def _double_mean(a, b):
    """double-precision mean SSIM of two contiguous planes by the strongest checker present: the reference's own naive
    template when it was compiled here, else our C restatement with true-math taps"""
    if oracle.have_naive():
        return oracle.naive_ssim(a, b)[0]
    _, total, _ = oracle.oracle_ssim(a, b, taps=oracle.TAPS_RUNTIME)
    return total / a.size


# ------------------------------------------------------------------------------------------------ CPU: the reader
def test_reader_reproduces_the_reference_known_answers_360p(decoded, bbb360_full, golden):
    ref = golden["bbb_reference"]
    png = bbb360_full["png"]
    assert sorted(decoded) == ref["qualities"]
    assert (decoded[50] == bbb360_full["jpg50"]).all()                      # the committed decoded fixture is this reader's output
    worst = 0.0
    for suite, (w, h) in CROPS.items():
        for qi, q in enumerate(ref["qualities"]):
            assert decoded[q].shape == (360, 640, 3)
            for ch in range(3):
                mean = _double_mean(np.ascontiguousarray(png[:h, :w, ch]), np.ascontiguousarray(decoded[q][:h, :w, ch]))
                d = abs(mean - float(ref["known_answers"][suite][qi][ch]))
                worst = max(worst, d)
                assert d <= REF_TOLERANCE, (suite, q, ch, mean)
    print("99 known answers, worst |d| = %.1e" % worst)


@pytest.mark.skipif(not os.path.isdir(REF_IMAGES), reason="the 1080p JPEG files live in /root/reference only")
@pytest.mark.parametrize("quality", [0, 50, 100])
def test_reader_reproduces_the_reference_known_answers_1080p(quality, golden, bbb1080_green):
    from PIL import Image
    ref = golden["bbb_reference"]
    png = np.asarray(Image.open(os.path.join(REF_IMAGES, "big_buck_bunny_1080_07806.png")).convert("RGB"), dtype=np.uint8)
    with open(os.path.join(REF_IMAGES, "big_buck_bunny_1080_07806_%02d.jpg" % quality), "rb") as fh:
        jpg = api.decode_jpeg(fh.read())
    assert jpg.shape == (1080, 1920, 3)
    if quality == 50:
        assert (jpg[..., 1] == bbb1080_green["jpg50"]).all() and (png[..., 1] == bbb1080_green["png"]).all()
    for ch in range(3):
        _, total, _ = oracle.oracle_ssim(np.ascontiguousarray(png[..., ch]), np.ascontiguousarray(jpg[..., ch]), taps=oracle.TAPS_RUNTIME)
        assert abs(total / png[..., ch].size - float(ref["known_answers"]["bbb1080"][quality // 10][ch])) <= REF_TOLERANCE


def test_oracle_restatement_on_the_reference_known_answers(decoded, bbb360_full, golden):
    """our own C restatement (not the reference's template) against the same constants: the oracle pinned on real content"""
    ref = golden["bbb_reference"]
    png = bbb360_full["png"]
    for qi, q in enumerate(ref["qualities"]):
        for ch in range(3):
            _, total, _ = oracle.oracle_ssim(np.ascontiguousarray(png[..., ch]), np.ascontiguousarray(decoded[q][..., ch]), taps=oracle.TAPS_RUNTIME)
            assert abs(total / (640 * 360) - float(ref["known_answers"]["bbb360"][qi][ch])) <= REF_TOLERANCE


def _encode(img, **kw):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def _pil(data, mode):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert(mode), dtype=np.uint8)


@pytest.mark.parametrize("kw", [dict(quality=85, subsampling=0), dict(quality=85, subsampling=0, progressive=True),
                                dict(quality=60, subsampling=2), dict(quality=60, subsampling=2, progressive=True),
                                dict(quality=75, subsampling=0, optimize=True), dict(quality=30, subsampling=0, progressive=True, optimize=True)])
def test_reader_against_libjpeg_on_other_variants(kw, bbb360_full):
    """baseline / progressive, 4:4:4 / 4:2:0, default and optimised Huffman tables, odd sizes (partial MCUs), gray: the
    entropy decoding is exact, so what is left against libjpeg (PIL) is its different IDCT / colour rounding"""
    pytest.importorskip("PIL")
    for (h, w) in [(360, 640), (37, 53), (8, 8), (1, 1), (17, 300)]:
        src = np.ascontiguousarray(bbb360_full["png"][:h, :w])
        data = _encode(src, **kw)
        mine, theirs = api.decode_jpeg(data), _pil(data, "RGB")
        assert mine.shape == theirs.shape == (h, w, 3)
        d = np.abs(mine.astype(int) - theirs.astype(int))
        assert d.max() <= 4 and d.mean() <= 0.15, (kw, h, w, d.max(), d.mean())
        gray = np.ascontiguousarray(src[..., 1])
        data = _encode(gray, **{k: v for k, v in kw.items() if k != "subsampling"})
        mine, theirs = api.decode_jpeg(data), _pil(data, "L")
        assert mine.shape == theirs.shape == (h, w)
        assert np.abs(mine.astype(int) - theirs.astype(int)).max() <= 1


def test_reader_restart_intervals(bbb360_full):
    """DRI / RSTn: the same coefficients with and without restart markers must decode to the same pixels"""
    pytest.importorskip("PIL")
    src = np.ascontiguousarray(bbb360_full["png"][:100, :200])
    plain = _encode(src, quality=80, subsampling=0)
    try:
        marked = _encode(src, quality=80, subsampling=0, restart_marker_blocks=5)
    except TypeError:
        pytest.skip("this PIL cannot write restart markers")
    if b"\xff\xdd" not in marked:
        pytest.skip("this PIL ignored restart_marker_blocks")
    assert (api.decode_jpeg(plain) == api.decode_jpeg(marked)).all()


def test_reader_rejects_what_it_cannot_decode(jpeg_files):
    good = jpeg_files["q50"].tobytes()
    for bad in (b"", b"\xff\xd8", b"\x89PNG\r\n\x1a\n" + b"0" * 32, good[:2] + b"\xff\xe0\x00\x01", good[:400],
                good.replace(b"\xff\xc2", b"\xff\xc9", 1),          # arithmetic coding
                good.replace(b"\xff\xc4", b"\xff\xfe")):            # every Huffman table turned into a comment
        with pytest.raises(ValueError):
            api.decode_jpeg(bad)
    # a truncated entropy-coded segment still has its header: the decoder must not crash or read out of bounds
    for cut in (len(good) // 2, len(good) - 3):
        try:
            img = api.decode_jpeg(good[:cut] + b"\xff\xd9")
            assert img.shape == (360, 640, 3)
        except ValueError:
            pass
    lib = api.imgio_lib()
    import ctypes as C
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    buf = np.frombuffer(good, dtype=np.uint8)
    small = np.empty(16, dtype=np.uint8)
    import errno
    assert lib.ssim_imgio_decode_jpeg(buf.ctypes.data, buf.size, small.ctypes.data, small.size, C.byref(w), C.byref(h), C.byref(c)) == errno.ERANGE
    assert (w.value, h.value, c.value) == (640, 360, 3) and b"too small" in lib.ssim_imgio_last_error()


def test_reader_survives_damaged_files(jpeg_files):
    """a few hundred mutants (bit flips, stray 0xFF, truncations, damaged headers) of a progressive file and of a baseline
    4:2:0 one: every call returns -- pixels or ValueError -- and never crashes the process (tools/dev/jpeg_fuzz.cpp is the
    same loop under AddressSanitizer / UBSan)"""
    rng = np.random.default_rng(20261017)
    bases = [jpeg_files["q10"].copy()]
    try:
        bases.append(np.frombuffer(_encode(api.decode_jpeg(jpeg_files["q90"])[:96, :160], quality=70, subsampling=2), dtype=np.uint8).copy())
    except ImportError:
        pass
    decoded = 0
    for base in bases:
        for _ in range(150):
            d = base.copy()
            for _ in range(int(rng.integers(1, 8))):
                pos = int(rng.integers(0, min(d.size, 700) if rng.integers(0, 4) == 0 else d.size))
                kind = int(rng.integers(0, 4))
                if kind == 0:
                    d[pos] = rng.integers(0, 256)
                elif kind == 1:
                    d[pos] ^= 1 << int(rng.integers(0, 8))
                elif kind == 2:
                    d[pos] = 0xFF
                elif d.size > 10:
                    d = d[:pos + 1].copy()
            try:
                img = api.decode_jpeg(d)
                assert img.ndim in (2, 3) and img.size > 0
                decoded += 1
            except ValueError:
                pass
    assert decoded > 0


def test_cli_reads_jpeg(jpeg_files, decoded, tmp_path):
    from test_cli import fnv
    path = str(tmp_path / "q30.jpg")
    jpeg_files["q30"].tofile(path)
    out = subprocess.run([CLI, "--probe", path], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["640", "360", "3", fnv(decoded[30])]


# ------------------------------------------------------------------------------------------------ GPU: the reference's test
@pytest.mark.gpu
@pytest.mark.parametrize("suite", sorted(CROPS))
def test_gpu_bbb_suites_against_the_reference_known_answers(suite, decoded, bbb360_full, golden):
    """test_compute_ssim() of the reference (tests/rmgr-ssim-tests.cpp:228-335) for its 360p suites: images stay interleaved
    (init_interleaved: step = channels, stride = full frame), width / height are the crop, one call per channel with the map."""
    ref = golden["bbb_reference"]
    png = bbb360_full["png"]
    w, h = CROPS[suite]
    worst_g = worst_m = 0.0
    for qi, q in enumerate(ref["qualities"]):
        jpg = decoded[q]
        for ch in range(3):
            kw = dict(width=w, height=h, step_a=3, step_b=3, stride_a=640 * 3, stride_b=640 * 3, a_off=ch, b_off=ch)
            s, m = api.compute_ssim(png, jpg, want_map=True, **kw)                      # imgA = reference frame, imgB = JPEG, as there
            expected = float(ref["known_answers"][suite][qi][ch])
            worst_g = max(worst_g, abs(float(s) - expected))
            assert abs(float(s) - expected) <= GLOBAL_TOL, (suite, q, ch, s, expected)
            if q in (0, 50, 100):                                                       # per-pixel check (the oracle map costs CPU time)
                _, _, om = oracle.oracle_ssim(png, jpg, want_map=True, taps=oracle.TAPS_RUNTIME, **kw)
                worst_m = max(worst_m, float(np.abs(m - om).max()))
                assert np.abs(m - om).max() <= PIXEL_TOL
            s2, _ = api.compute_ssim(png, jpg, want_map=False, **kw)
            assert s2 == s
    print("%s: 33 known answers, worst |global - reference constant| = %.2e, worst map |d| = %.2e" % (suite, worst_g, worst_m))


@pytest.mark.gpu
def test_gpu_bbb1080_green_against_the_reference_known_answer(bbb1080_green, golden):
    s, _ = api.compute_ssim(bbb1080_green["png"], bbb1080_green["jpg50"])
    assert abs(float(s) - float(golden["bbb_reference"]["known_answers"]["bbb1080"][5][1])) <= GLOBAL_TOL


@pytest.mark.gpu
def test_cli_on_a_png_jpeg_pair(jpeg_files, decoded, bbb360_full, golden, tmp_path):
    """what a user of the reference's CLI does with its test images: rmgr-ssim frame.png frame_50.jpg (src/ssim-cli.cpp:216-389)"""
    from test_cli import write_png
    a, b = str(tmp_path / "frame.png"), str(tmp_path / "frame_50.jpg")
    write_png(a, bbb360_full["png"])
    jpeg_files["q50"].tofile(b)
    r = subprocess.run([CLI, b, a], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = [float(x) for x in golden["bbb_reference"]["known_answers"]["bbb360"][5]]
    lines = r.stdout.splitlines()
    for ch in range(3):
        assert lines[ch] == "Channel %u: % 7.4f" % (ch, want[ch])
    assert lines[3] == "Average  : % 7.4f" % (sum(want) / 3)
