#!/usr/bin/env python
"""Regenerates tests/golden/*.{npz,json}.  Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_golden.py

Inputs : /root/reference/tests/images.  PNG decoded with PIL (lossless: decoder-independent); JPEG decoded with this
         repository's own reader (ssim_b200/csrc/jpeg_reader.h through libssim_imgio.so), whose pixels are identical to
         those of the decoder the reference's tests use -- checked below: the reference's hard-coded bbb known answers
         (tests/rmgr-ssim-tests.cpp:388-465, 132 numbers) are reproduced to 1e-13 by its own naive template on them.
Outputs: decoded pixels, the 360p JPEG files themselves (as bytes: the GPU box decodes them with the same reader), what the
         UNMODIFIED reference build (oracle/_ref) returns for them, and the reference's known answers.
The six einstein known answers are the reference's own constants (tests/rmgr-ssim-tests.cpp:354-359)."""
import re
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import ssim_b200.api as api  # noqa: E402
from ssim_b200.synth import checksum, synth_pair  # noqa: E402

IMAGES = "/root/reference/tests/images"
OUT = os.path.dirname(os.path.abspath(__file__))

# reference tests/rmgr-ssim-tests.cpp:341-359 (files and their 18-digit known answers, double mean)
EINSTEIN = [
    ("einstein", "1.000000000000000000000000000000000"),
    ("meanshift", "0.987345868581455342542598819456431"),
    ("contrast", "0.901217091012390185892926336265424"),
    ("impulse", "0.839533769204009687363862456348761"),
    ("blur", "0.702192033056262932311859850040160"),
    ("jpg", "0.669938383706498006524758818118705"),
]

EDGE_DIMS = [(1, 1), (2, 3), (7, 3), (5, 5), (11, 11), (16, 16), (255, 63), (256, 64), (257, 65), (300, 1),
             (1, 300), (513, 129), (64, 75), (65, 11), (130, 200)]


QUALITIES = list(range(0, 101, 10))


def f32(x):
    return float(np.float32(x))


def load_jpeg(stub, quality):
    with open(os.path.join(IMAGES, "%s_%02d.jpg" % (stub, quality)), "rb") as fh:
        data = fh.read()
    return data, api.decode_jpeg(data)


def reference_bbb_constants():
    """The 4 x 11 x 3 known answers of the reference's bbb suites, as the strings its test file holds
    (tests/rmgr-ssim-tests.cpp:388-465): suite -> [quality 0, 10, .. 100][channel]."""
    text = open("/root/reference/tests/rmgr-ssim-tests.cpp").read()
    out = {}
    for suite in ("bbb360", "bbb1080", "bbb255", "bbb257"):
        body = text[text.index("static void test_%s(" % suite):]
        body = body[:body.index("test_bbb(")]
        rows = re.findall(r"REF_SSIM3\(\s*([0-9.]+)\s*,\s*([0-9.]+)\s*,\s*([0-9.]+)\s*\)", body)
        assert len(rows) == 11, (suite, len(rows))
        out[suite] = [list(r) for r in rows]
    return out


def pin_decoder(consts):
    """Every one of the reference's 132 bbb known answers from ITS files, OUR reader and ITS naive template: |d| <= 1e-13
    (REF_TOLERANCE, tests/rmgr-ssim-tests.cpp:72).  A single differing pixel moves these means by ~1e-8."""
    worst = 0.0
    for suite, stub, (w, h) in (("bbb360", "big_buck_bunny_360_07806", (640, 360)), ("bbb1080", "big_buck_bunny_1080_07806", (1920, 1080)),
                                ("bbb255", "big_buck_bunny_360_07806", (255, 63)), ("bbb257", "big_buck_bunny_360_07806", (257, 65))):
        png = np.asarray(Image.open(os.path.join(IMAGES, stub + ".png")).convert("RGB"), dtype=np.uint8)
        for qi, q in enumerate(QUALITIES):
            _, jpg = load_jpeg(stub, q)
            for ch in range(3):
                mean, _ = oracle.naive_ssim(np.ascontiguousarray(png[:h, :w, ch]), np.ascontiguousarray(jpg[:h, :w, ch]))
                d = abs(mean - float(consts[suite][qi][ch]))
                worst = max(worst, d)
                assert d <= 1e-13, (suite, q, ch, mean, consts[suite][qi][ch])
    print("JPEG reader pinned: 132 reference known answers reproduced, worst |d| = %.1e" % worst)
    return worst


def main():
    oracle.build()
    assert oracle.have_ref(), "needs /root/reference to build oracle/_ref"

    # ---- einstein: 256x256 8-bit grayscale (BASELINE.json configs[0])
    planes = {}
    for name, _ in EINSTEIN:
        img = Image.open(os.path.join(IMAGES, name + ".png"))
        assert img.mode == "L" and img.size == (256, 256), (name, img.mode, img.size)
        planes[name] = np.asarray(img, dtype=np.uint8).copy()
    np.savez_compressed(os.path.join(OUT, "einstein.npz"), **planes)
    ein = {}
    ref = planes["einstein"]
    for name, golden in EINSTEIN:
        a = planes[name]
        r64a, m64a = oracle.ref_ssim("f64", a, ref, want_map=True)
        r64g, _ = oracle.ref_ssim("f64", a, ref, impl=oracle.IMPL_GENERIC)
        r32, _ = oracle.ref_ssim("f32", a, ref, openmp=True)
        ein[name] = {"golden_double_mean": golden, "ref_f64_auto": f32(r64a), "ref_f64_generic": f32(r64g),
                     "ref_f32_auto_openmp": f32(r32),
                     "ref_f64_auto_map_min": f32(m64a.min()), "ref_f64_auto_map_sum64": float(m64a.astype(np.float64).sum())}
    # one full map as a fixture (blur: lowest-scoring smooth case)
    _, m = oracle.ref_ssim("f64", planes["blur"], ref, want_map=True)
    np.savez_compressed(os.path.join(OUT, "einstein_blur_map_f64auto.npz"), map=m)

    # ---- bbb 360p: RGB interleaved (step=3), crops 255x63 / 257x65 with the full-frame stride kept
    png = np.asarray(Image.open(os.path.join(IMAGES, "big_buck_bunny_360_07806.png")).convert("RGB"), dtype=np.uint8)
    _, jpg = load_jpeg("big_buck_bunny_360_07806", 50)
    assert png.shape == (360, 640, 3) and jpg.shape == png.shape
    np.savez_compressed(os.path.join(OUT, "bbb360.npz"), png=png, jpg50=jpg)       # the full frames; tests slice the top 80 rows
    # the eleven 360p JPEG files as bytes (0.7 MB): decoded on the spot by the tests with the same reader
    np.savez(os.path.join(OUT, "bbb360_jpeg_files.npz"),
             **{"q%02d" % q: np.frombuffer(load_jpeg("big_buck_bunny_360_07806", q)[0], dtype=np.uint8) for q in QUALITIES})
    consts = reference_bbb_constants()
    worst = pin_decoder(consts)
    bbb = {}
    # full frames, all three channels, and the 1080p frame (green channel only: fixture size), like the reference's
    # tests/rmgr-ssim-tests.cpp:388-425
    for ch in range(3):
        kw = dict(step_a=3, step_b=3, stride_a=640 * 3, stride_b=640 * 3, width=640, height=360, a_off=ch, b_off=ch)
        r64a, m64a = oracle.ref_ssim("f64", jpg, png, want_map=True, **kw)
        r32, _ = oracle.ref_ssim("f32", jpg, png, openmp=True, **kw)
        bbb["640x360_ch%d" % ch] = {"ref_f64_auto": f32(r64a), "ref_f32_auto_openmp": f32(r32), "ref_f64_auto_map_sum64": float(m64a.astype(np.float64).sum()),
                                    "ref_f64_auto_map_min": f32(m64a.min())}
    big_png = np.asarray(Image.open(os.path.join(IMAGES, "big_buck_bunny_1080_07806.png")).convert("RGB"), dtype=np.uint8)[..., 1].copy()
    big_jpg = load_jpeg("big_buck_bunny_1080_07806", 50)[1][..., 1].copy()
    assert big_png.shape == (1080, 1920)
    np.savez_compressed(os.path.join(OUT, "bbb1080_green.npz"), png=big_png, jpg50=big_jpg)
    r64a, m64a = oracle.ref_ssim("f64", big_jpg, big_png, want_map=True)
    r32, _ = oracle.ref_ssim("f32", big_jpg, big_png, openmp=True)
    bbb["1920x1080_green"] = {"ref_f64_auto": f32(r64a), "ref_f32_auto_openmp": f32(r32), "ref_f64_auto_map_sum64": float(m64a.astype(np.float64).sum()),
                              "ref_f64_auto_map_min": f32(m64a.min())}
    rows = 80  # the crops below use the first 80 rows only; stride stays 640*3
    png, jpg = png[:rows].copy(), jpg[:rows].copy()
    for (w, h) in [(255, 63), (257, 65), (640, 80)]:
        for ch in range(3):
            kw = dict(step_a=3, step_b=3, stride_a=640 * 3, stride_b=640 * 3, width=w, height=h, a_off=ch, b_off=ch)
            r64a, m64a = oracle.ref_ssim("f64", jpg, png, want_map=True, **kw)
            r64g, _ = oracle.ref_ssim("f64", jpg, png, impl=oracle.IMPL_GENERIC, **kw)
            r32, _ = oracle.ref_ssim("f32", jpg, png, **kw)
            bbb["%dx%d_ch%d" % (w, h, ch)] = {"ref_f64_auto": f32(r64a), "ref_f64_generic": f32(r64g), "ref_f32_auto": f32(r32),
                                             "ref_f64_auto_map_sum64": float(m64a.astype(np.float64).sum())}

    # ---- synthetic recipe, seed 0x5517
    syn = {}
    for (w, h, f) in [(1920, 1080, 0), (1920, 1080, 1), (1920, 1080, 4095), (3840, 2160, 0)]:
        a, b = synth_pair(w, h, f)
        r64a, _ = oracle.ref_ssim("f64", a, b, openmp=True)
        r64g, _ = oracle.ref_ssim("f64", a, b, impl=oracle.IMPL_GENERIC, openmp=True)
        r32, _ = oracle.ref_ssim("f32", a, b, openmp=True)
        syn["%dx%d_f%d" % (w, h, f)] = {"checksum": "%016x" % checksum(a, b), "ref_f64_auto": f32(r64a),
                                        "ref_f64_generic": f32(r64g), "ref_f32_auto_openmp": f32(r32)}
    a, b = synth_pair(16384, 16384, 0)
    r64a, _ = oracle.ref_ssim("f64", a, b, openmp=True)
    r32, _ = oracle.ref_ssim("f32", a, b, openmp=True)
    syn["16384x16384_f0"] = {"ref_f64_auto": f32(r64a), "ref_f32_auto_openmp": f32(r32)}
    for (w, h) in EDGE_DIMS:
        a, b = synth_pair(w, h, 3)
        r64a, m64a = oracle.ref_ssim("f64", a, b, want_map=True)
        r64g, _ = oracle.ref_ssim("f64", a, b, impl=oracle.IMPL_GENERIC)
        r32, _ = oracle.ref_ssim("f32", a, b)
        syn["%dx%d_f3" % (w, h)] = {"ref_f64_auto": f32(r64a), "ref_f64_generic": f32(r64g), "ref_f32_auto": f32(r32),
                                    "ref_f64_auto_map_sum64": float(m64a.astype(np.float64).sum())}

    # ---- 16-bit (L = 65535): the reference's own naive::compute_ssim<double, uint16_t> (tests/ssim_naive.h:230-339,
    #      compiled in place into oracle/_ref/libnaive.so) on genuinely 16-bit inputs
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import u16_inputs  # noqa: E402
    assert oracle.have_naive()
    u16 = {}
    for (h, w) in u16_inputs.U16_SHAPES:
        a, b = u16_inputs.pair16(h, w, u16_inputs.seed_of(h, w))
        mean, m = oracle.naive_ssim(a, b, want_map=True)
        u16["%dx%d" % (w, h)] = {"inputs_sha256_16": u16_inputs.digest(a, b), "naive_double_mean": repr(mean),
                                 "naive_map_sum": repr(float(m.sum())), "naive_map_min": repr(float(m.min()))}
    h, w = u16_inputs.U16_FIXTURE_SHAPE
    a, b = u16_inputs.pair16(h, w, 2024)
    mean, m = oracle.naive_ssim(a, b, want_map=True)
    np.savez_compressed(os.path.join(OUT, "u16_pair.npz"), a=a, b=b, naive_map=m.astype(np.float32))
    u16["fixture_%dx%d" % (w, h)] = {"naive_double_mean": repr(mean)}
    # the same shim, 8-bit instantiation, must reproduce the reference's einstein known answers (they were made with it)
    for name, golden in EINSTEIN:
        mean, _ = oracle.naive_ssim(planes[name], ref)
        assert abs(mean - float(golden)) < 1e-13, (name, mean, golden)

    with open(os.path.join(OUT, "golden.json"), "w") as fh:
        json.dump({"einstein": ein, "bbb360_jpg50": bbb, "synthetic": syn, "u16_naive": u16,
                   "bbb_reference": {"known_answers": consts, "qualities": QUALITIES, "decoder_pin_worst_abs_diff": worst,
                                     "source": "tests/rmgr-ssim-tests.cpp:388-465 of the reference (double means of its naive template); "
                                               "suite -> [quality][channel]; bbb255 / bbb257 are the 255x63 / 257x65 crops of the 360p frame"},
                   "note": "ref_* values are float32 results of the unmodified reference build (oracle/_ref), repr as double"},
                  fh, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
