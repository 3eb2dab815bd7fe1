"""CPU-only tests of the host side: ABI layouts, exported symbols, pure-host API helpers, sharding logic
(world_size-2 gloo).  No compute call is made here: there is no GPU in the build container and no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
from ssim_b200 import _abi, api, parallel
from ssim_b200.synth import synth_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_struct_layouts_match_reference_abi():
    """24 / 96 / 24 bytes on LP64, field offsets as in the reference's ssim.h:489-533"""
    assert C.sizeof(_abi.ImgParams) == 24 and C.sizeof(_abi.Params) == 96 and C.sizeof(_abi.ThreadPool) == 24
    assert _abi.Params.imgA.offset == 8 and _abi.Params.imgB.offset == 32 and _abi.Params.ssimMap.offset == 56
    assert _abi.Params.ssimStep.offset == 64 and _abi.Params.ssimStride.offset == 72 and _abi.Params.alloc.offset == 80


def test_libssim_cuda_exports_every_declared_symbol():
    lib = api.cuda_lib()
    names = _declared_symbols("ssim_cuda.h", "ssim_cuda_")
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ssim_cuda_abi_version() == 2
    assert lib.ssim_cuda_last_error_string() is not None


def test_librmgr_exports_reference_api():
    lib = api.rmgr_lib()
    for n in _declared_symbols("rmgr/ssim.h", "rmgr_ssim_") + _declared_symbols("rmgr/ssim-openmp.h", "rmgr_ssim_"):
        if n.endswith("Fct"):
            continue
        assert hasattr(lib, n), n
    # the two C++ overloads (mangled exactly as a caller compiled against the reference's header expects)
    out = subprocess.run(["nm", "-D", os.path.join(api.LIB_DIR, "librmgr-ssim.so")], capture_output=True, text=True).stdout
    assert "_ZN4rmgr4ssim12compute_ssimEPfRK17rmgr_ssim_Params_PK21rmgr_ssim_ThreadPool_" in out
    assert "_ZN4rmgr4ssim12compute_ssimERKNS0_6ParamsE" in out


def test_pure_host_api_helpers():
    """get_version / init_interleaved / init_planar / use_default_allocator: reference src/ssim.cpp:156-217,1126-1142"""
    lib = api.rmgr_lib()
    assert api.get_version() == (2, 1, 0, "2.1.0")
    assert lib.rmgr_ssim_get_version(None) == 22
    buf = np.zeros(64, np.uint8)
    ip = _abi.ImgParams()
    assert lib.rmgr_ssim_init_interleaved(C.byref(ip), buf.ctypes.data, -24, 3, 2) == 0
    assert (ip.topLeft, ip.step, ip.stride) == (buf.ctypes.data + 2, 3, -24)
    assert lib.rmgr_ssim_init_interleaved(C.byref(ip), buf.ctypes.data, 24, 3, 3) == 22       # channelNum >= channelCount
    assert lib.rmgr_ssim_init_interleaved(None, buf.ctypes.data, 24, 3, 0) == 22
    assert lib.rmgr_ssim_init_interleaved(C.byref(ip), None, 24, 3, 0) == 22
    planes = (C.c_void_p * 2)(buf.ctypes.data, buf.ctypes.data + 32)
    strides = (C.c_ssize_t * 2)(8, 16)
    assert lib.rmgr_ssim_init_planar(C.byref(ip), planes, strides, 1) == 0
    assert (ip.topLeft, ip.step, ip.stride) == (buf.ctypes.data + 32, 1, 16)
    assert lib.rmgr_ssim_init_planar(C.byref(ip), None, strides, 0) == 22
    p = _abi.Params()
    assert lib.rmgr_ssim_use_default_allocator(C.byref(p)) == 0 and p.alloc and p.dealloc
    assert lib.rmgr_ssim_use_default_allocator(None) == 22


def test_validation_happens_before_any_device_work():
    """EINVAL paths of compute_ssim (reference src/ssim.cpp:962-978) return without touching CUDA"""
    lib = api.rmgr_lib()
    a = np.zeros((4, 4), np.uint8)
    out = C.c_float()
    p = _abi.make_params(a, a, 4, 4)
    assert lib.rmgr_ssim_compute_ssim(None, C.byref(p), None) == 22
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), None, None) == 22
    p.imgB.topLeft = None
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 22
    p = _abi.make_params(a, a, 4, 0)
    assert lib.rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 22


def test_no_cpu_fallback_without_a_device():
    lib = api.cuda_lib()
    if lib.ssim_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    a = np.zeros((8, 8), np.uint8)
    with pytest.raises(api.SsimError) as e:
        api.compute_ssim(a, a)
    assert e.value.errno == 19                                                              # ENODEV, never a silent CPU result


def test_strip_bounds_cover_the_image():
    for h, n in [(16384, 8), (2160, 4), (7, 8), (1080, 3), (11, 2)]:
        rows = 0
        for g in range(n):
            s0, s1, oy, orows = parallel.strip_bounds(h, n, g)
            assert 0 <= s0 <= s1 <= h and s0 + oy + orows <= s1
            assert (s0 + oy) == h * g // n
            # interior edges carry the full halo, exterior edges none
            assert oy == min(5, h * g // n) and s1 - (s0 + oy + orows) == min(5, h - h * (g + 1) // n)
            rows += orows
        assert rows == h
    assert [parallel.shard_frames(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import oracle
from ssim_b200 import parallel
from ssim_b200.synth import synth_pair
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
W, H = 150, 61
# --- one image in strips with halos: each rank only ever touches ITS rows; partial sums all-reduced (SUM of doubles)
s0, s1, oy, orows = parallel.strip_bounds(H, world, rank)
a, b = synth_pair(W, s1 - s0, 4, y0=s0)
_, _, m = oracle.oracle_ssim(a, b, want_map=True)            # clamps at the strip's own edges
part = torch.tensor([float(m[oy:oy + orows].astype(np.float64).sum())], dtype=torch.float64)
dist.all_reduce(part, op=dist.ReduceOp.SUM)
fa, fb = synth_pair(W, H, 4)
full, tot, fm = oracle.oracle_ssim(fa, fb, want_map=True)
assert abs(part.item() - fm.astype(np.float64).sum()) < 1e-9, (part.item(), tot)
assert parallel.mean_from_partials(part.item(), W, H) == np.float32(fm.astype(np.float64).sum() / (W * H))
assert np.abs(m[oy:oy + orows] - fm[s0 + oy:s0 + oy + orows]).max() == 0.0     # halo rows make strips exact
# --- a frame batch sharded across ranks, results gathered
f0, f1 = parallel.shard_frames(5, world, rank)
mine = torch.zeros(5, dtype=torch.float64)
for f in range(f0, f1):
    x, y = synth_pair(40, 30, f)
    mine[f] = float(oracle.oracle_ssim(x, y)[0])
dist.all_reduce(mine, op=dist.ReduceOp.SUM)
want = [float(oracle.oracle_ssim(*synth_pair(40, 30, f))[0]) for f in range(5)]
assert np.allclose(mine.numpy(), want, atol=0), (mine, want)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gloo_strips_and_batch(tmp_path):
    """N>1 host logic on CPU: world_size 2, gloo backend, rendezvous on 127.0.0.1"""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    procs = [subprocess.Popen([sys.executable, str(script), str(port), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


@pytest.mark.parametrize("name", ["fast_div_check", "plan_check"])
def test_host_side_kernel_helpers(tmp_path, name):
    """Host-only logic that lives in ssim_kernels.h, compiled as plain C++: fast_div() (the kernel's work-item decode) against
    plain division, plan_slots() + the PieceCursor (the persistent kernel's work partition) for coverage invariants and the documented plans."""
    exe = str(tmp_path / name)
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clients", name + ".cpp")
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ssim_b200", "csrc")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-I", inc, "-I", "/usr/local/cuda/include", src, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and " ok" in r.stdout, r.stdout


def test_hot_code_of_the_fused_kernel_fits_the_instruction_cache():
    """The two hot loops of every variant of the fused kernel must stay within 32 KB of addresses ("code
    layout"): beyond that the kernel loses ~9% to instruction fetch.  Checked on the built library with cuobjdump (no GPU)."""
    import re
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_summary.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sizes = [float(x) for x in re.findall(r"hot code: .* = ([0-9.]+) KB", r.stdout)]
    assert len(sizes) == 5, r.stdout                      # five instantiations of ssim_fused_kernel
    assert max(sizes) < 32.0, sizes
    assert r.stdout.count("no local-memory spills") >= 5, r.stdout       # and none of them spills


def _byte_perm(x, y, sel):
    """CUDA's __byte_perm / PTX prmt.b32 (default mode): result byte k = byte (nibble k of sel) of the 8-byte pool {y, x}."""
    pool = [(x >> (8 * i)) & 0xff for i in range(4)] + [(y >> (8 * i)) & 0xff for i in range(4)]
    return sum(pool[(sel >> (4 * k)) & 7] << (8 * k) for k in range(4))


def test_pixel_widening_tricks_are_exact():
    """The producer's pixel -> f32 conversions (ssim_kernels.cu, "Widening") restated on the CPU.  8-bit: one PRMT with the magic
    word 0x4B000064 puts two pixels under the f16 exponent byte 0x64 (= 1024 + pixel each), the mixed-precision add
    f16 + f32 -> f32 with -(1024 + centre) leaves pixel - centre; the 26 columns a lane needs are bytes 11..36 of its 48-byte
    window, taken as one single, twelve pairs, one single.  16-bit: selectors 0x7610 / 0x7632 put a pixel under 0x4B00 (2^23 + pixel
    as f32).  Everything must be exact for every pixel / centre value."""
    import numpy as np
    magic = 0x4B000064
    # f16 bit pattern 0x64pp is 1024 + pp, and the add is exact in f32 for every pixel and centre
    p = np.arange(256, dtype=np.uint16)
    h = (np.uint16(0x6400) | p).view(np.float16)
    assert np.array_equal(h.astype(np.float32), 1024.0 + p.astype(np.float32))
    for c in range(256):
        neg = np.float32(-(1024.0 + c))
        assert np.array_equal(h.astype(np.float32) + neg, p.astype(np.float32) - np.float32(c))
    # the pairing schedule of the kernel's column loop over a window of 48 bytes held in 12 words
    rng = np.random.default_rng(5)
    window = rng.integers(0, 256, 48, dtype=np.uint8)
    words = [int.from_bytes(window[4 * i:4 * i + 4].tobytes(), "little") for i in range(12)]
    got, hp = [], 0
    for ii in range(26):
        byte_idx = ii + 11
        if ii == 0 or (ii & 1):
            sel = 0x4343 if (byte_idx & 3) == 3 else 0x4342 if (byte_idx & 2) else 0x4140
            hp = _byte_perm(words[byte_idx >> 2], magic, sel)
            half = hp & 0xffff
        else:
            half = hp >> 16
        assert half >> 8 == 0x64
        got.append(half & 0xff)
    assert got == [int(v) for v in window[11:37]]
    # 16-bit pixels: 2^23 + pixel as f32 bits
    w = 0xBEEF1234
    lo, hi = _byte_perm(w, magic, 0x7610), _byte_perm(w, magic, 0x7632)
    assert np.array([lo, hi], dtype=np.uint32).view(np.float32).tolist() == [8388608.0 + 0x1234, 8388608.0 + 0xBEEF]
