"""Source-level drop-in: callers written against the reference's C and C++ API compile (C99, C++98 and C++17, -Wall -Wextra
-Werror) against include/rmgr/*.h, link with librmgr-ssim.so and behave like the reference (validation on the CPU, results
on the GPU)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GLOBAL_TOL
from ssim_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INC = os.path.join(ROOT, "include")


def _build(tmp, src, std, cxx):
    exe = os.path.join(tmp, os.path.basename(src) + "." + std.replace("+", "p"))
    cmd = ["/usr/bin/g++" if cxx else "/usr/bin/gcc", "-std=" + std, "-O1", "-Wall", "-Wextra", "-Werror", "-I", INC, os.path.join(HERE, "clients", src),
           "-o", exe, "-L", api.LIB_DIR, "-lrmgr-ssim", "-lssim_cuda", "-Wl,-rpath," + api.LIB_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.fixture(scope="module")
def clients(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("clients"))
    return {"c99": _build(tmp, "client.c", "c99", False), "cpp98": _build(tmp, "client.cpp", "c++98", True),
            "cpp17": _build(tmp, "client.cpp", "c++17", True), "dir": tmp}


def test_clients_compile_and_validate(clients):
    for k in ("c99", "cpp98", "cpp17"):
        r = subprocess.run([clients[k]], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.startswith("validation ok"), (k, r.returncode, r.stdout, r.stderr)


def _write_pair(path, a, b):
    h, w = a.shape[:2]
    c = 1 if a.ndim == 2 else a.shape[2]
    with open(path, "wb") as f:
        f.write(b"%d %d %d\n" % (w, h, c))
        f.write(np.ascontiguousarray(a).tobytes())
        f.write(np.ascontiguousarray(b).tobytes())


@pytest.mark.gpu
def test_clients_results(clients, einstein, golden, bbb360):
    # gray pair through the C and both C++ builds: the reference's known answer, all call forms agree
    path = os.path.join(clients["dir"], "gray.bin")
    _write_pair(path, einstein["einstein"], einstein["blur"])
    want = float(golden["einstein"]["blur"]["golden_double_mean"])
    r = subprocess.run([clients["c99"], path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert all(abs(float(v) - want) <= GLOBAL_TOL for v in r.stdout.split())
    for k in ("cpp98", "cpp17"):
        r = subprocess.run([clients[k], path], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        s1, s2, s3, msum, s4 = (float(v) for v in r.stdout.split())
        assert abs(s1 - want) <= GLOBAL_TOL and s1 == s2 == s3 == s4
        assert abs(msum / (256 * 256) - s1) <= 1e-6
    # interleaved RGB (step = 3), one call per channel like tests/rmgr-ssim-tests.cpp:273-300
    from oracle import oracle_ssim
    a, b = bbb360["png"][:90, :160], bbb360["jpg50"][:90, :160]
    path = os.path.join(clients["dir"], "rgb.bin")
    _write_pair(path, a, b)
    r = subprocess.run([clients["cpp17"], path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 3
    for ch, line in enumerate(lines):
        o, _, _ = oracle_ssim(np.ascontiguousarray(a[..., ch]), np.ascontiguousarray(b[..., ch]))
        assert abs(float(line.split()[0]) - float(o)) <= GLOBAL_TOL
