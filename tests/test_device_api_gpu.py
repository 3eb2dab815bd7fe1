"""GPU tests of the device-resident entry points (frames, strips, multi-GPU) against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle
from conftest import GLOBAL_TOL, PIXEL_TOL
from ssim_b200 import api, parallel
from ssim_b200.synth import synth_pair

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(x):
    return torch.from_numpy(x).cuda()


def test_batch_of_frames_matches_per_frame_oracle():
    """BASELINE.json configs[4] in miniature: N frame pairs, one launch, per-frame SSIM + maps"""
    F, W, H = 5, 208, 77                       # pitch 208 is a multiple of 16
    a = np.stack([synth_pair(W, H, f)[0] for f in range(F)])
    b = np.stack([synth_pair(W, H, f)[1] for f in range(F)])
    dA, dB = _dev(a), _dev(b)
    dMap = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
    dSums = torch.empty(F, dtype=torch.float64, device="cuda")
    dSsim = torch.empty(F, dtype=torch.float32, device="cuda")
    api.compute_device(0, torch.cuda.current_stream().cuda_stream, W, H, 0, H, F, dA.data_ptr(), W, W * H, dB.data_ptr(), W, W * H,
                       dMap.data_ptr(), W, W * H, dSums.data_ptr(), dSsim.data_ptr())
    torch.cuda.synchronize()
    assert api.cuda_lib().ssim_cuda_last_launch_count() == 1
    for f in range(F):
        o, tot, om = oracle.oracle_ssim(a[f], b[f], want_map=True)
        assert abs(float(dSsim[f]) - float(o)) <= GLOBAL_TOL
        assert abs(float(dSums[f]) - tot) <= 1e-6 * W * H
        assert np.abs(dMap[f].cpu().numpy() - om).max() <= PIXEL_TOL
    # synth_fill on the device produces the same bytes as the host recipe
    dA2, dB2 = torch.empty_like(dA[0]), torch.empty_like(dB[0])
    api.synth_fill(0, torch.cuda.current_stream().cuda_stream, dA2.data_ptr(), W, dB2.data_ptr(), W, W, H, 0, 3)
    assert np.array_equal(dA2.cpu().numpy(), a[3]) and np.array_equal(dB2.cpu().numpy(), b[3])


@pytest.mark.parametrize("world", [2, 3, 8])
def test_strips_with_halo_rows_reproduce_the_full_image(world):
    """BASELINE.json configs[3] in miniature: strips + 5 halo rows; sum of partial sums == full-image sum.
    All 'ranks' run on GPU 0 here; tests/test_host_logic.py covers the 2-process reduction with gloo."""
    W, H = 336, 203
    a, b = synth_pair(W, H, 6)
    o, tot, om = oracle.oracle_ssim(a, b, want_map=True)
    got_map = np.zeros((H, W), np.float32)
    total = 0.0
    for r in range(world):
        s0, s1, oy, orows = parallel.strip_bounds(H, world, r)
        dA, dB = _dev(np.ascontiguousarray(a[s0:s1])), _dev(np.ascontiguousarray(b[s0:s1]))
        dMap = torch.empty((orows, W), dtype=torch.float32, device="cuda")
        dSum = torch.empty(1, dtype=torch.float64, device="cuda")
        api.compute_device(0, None, W, s1 - s0, oy, orows, 1, dA.data_ptr(), W, 0, dB.data_ptr(), W, 0, dMap.data_ptr(), W, 0,
                           dSum.data_ptr(), None)
        torch.cuda.synchronize()
        got_map[s0 + oy:s0 + oy + orows] = dMap.cpu().numpy()
        total += float(dSum.item())
    assert abs(float(parallel.mean_from_partials(total, W, H)) - float(o)) <= GLOBAL_TOL
    assert np.abs(got_map - om).max() <= PIXEL_TOL


@pytest.mark.parametrize("geom", [(7, 5, 2), (6, 5, 1), (7, 0, 7), (3, 1, 2), (9, 5, 4), (12, 5, 2)])
def test_short_strips_clamp_rows_like_the_reference(geom):
    """Strips of fewer than 8 source rows (a 10-row image over 8 GPUs has them): rows below the strip replicate its last
    row (src/ssim.cpp:560-570), the TMA box being taller than the plane must not leak zero-filled rows."""
    src_rows, oy, orows = geom
    W = 200
    a, b = synth_pair(W, src_rows, 9)
    # the oracle on the same rows treats them as a whole image: its rows [oy, oy+orows) see the same clamped neighbours
    o, tot, om = oracle.oracle_ssim(a, b, want_map=True)
    pitch = 208
    dA = torch.zeros((src_rows, pitch), dtype=torch.uint8, device="cuda"); dB = torch.zeros_like(dA)
    dA[:, :W] = _dev(a); dB[:, :W] = _dev(b)
    dMap = torch.zeros((orows, W), dtype=torch.float32, device="cuda")
    dSum = torch.zeros(1, dtype=torch.float64, device="cuda")
    api.compute_device(0, None, W, src_rows, oy, orows, 1, dA.data_ptr(), pitch, 0, dB.data_ptr(), pitch, 0, dMap.data_ptr(), W, 0, dSum.data_ptr(), None)
    torch.cuda.synchronize()
    assert np.abs(dMap.cpu().numpy() - om[oy:oy + orows]).max() <= PIXEL_TOL
    assert abs(float(dSum.item()) - float(om[oy:oy + orows].astype(np.float64).sum())) <= 1e-4 * W * orows


@pytest.mark.parametrize("shape", [(37, 208, 77), (64, 64, 64), (130, 1280, 33), (300, 48, 20)])
def test_many_frames_are_reduced_per_frame(shape):
    """More frames than warp pairs, frames smaller than a slot's share, slots spanning several frames: every frame's sum must
    come out of the in-kernel per-frame reduction exactly once (and the accumulator words must be left clean for the next
    launch, which the second round checks)."""
    F, W, H = shape
    a = np.stack([synth_pair(W, H, f)[0] for f in range(F)])
    b = np.stack([synth_pair(W, H, f)[1] for f in range(F)])
    dA, dB = _dev(a), _dev(b)
    dSums = torch.zeros(F, dtype=torch.float64, device="cuda")
    dSsim = torch.zeros(F, dtype=torch.float32, device="cuda")
    dMap = torch.zeros((F, H, W), dtype=torch.float32, device="cuda")
    for rnd in range(2):
        dSums.zero_(); dSsim.zero_()
        api.compute_device(0, torch.cuda.current_stream().cuda_stream, W, H, 0, H, F, dA.data_ptr(), W, W * H, dB.data_ptr(), W, W * H,
                           dMap.data_ptr() if rnd == 0 else None, W, W * H, dSums.data_ptr(), dSsim.data_ptr())
        torch.cuda.synchronize()
        maps = dMap.cpu().numpy().astype(np.float64).sum(axis=(1, 2))
        assert np.abs(dSums.cpu().numpy() - maps).max() <= 1e-5 * W * H          # sums == sums of the stored maps
        for f in (0, 1, F // 2, F - 1):
            o, tot, _ = oracle.oracle_ssim(a[f], b[f])
            assert abs(float(dSsim[f]) - float(o)) <= GLOBAL_TOL, (f, rnd)


@pytest.mark.parametrize("shape", [(1, 208, 77), (9, 208, 77), (1, 1920, 270), (70, 64, 40)])
def test_negative_sums_survive_the_fixed_point_reduction(shape):
    """The in-kernel reduction adds the slots' sums as biased fixed-point integers (one packed atomic per slot and frame):
    anti-correlated images give NEGATIVE per-slot sums, which the bias must carry through; a second launch on the same
    stream checks that the accumulator words were put back to zero."""
    F, W, H = shape
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, (F, H, W), dtype=np.uint8)
    b = (255 - a).astype(np.uint8)                    # local structure inverted: SSIM < 0 almost everywhere
    b[F // 2] = a[F // 2]                             # ... and one frame that sums to exactly W*H
    dA, dB = _dev(a), _dev(b)
    dSums = torch.zeros(F, dtype=torch.float64, device="cuda")
    dSsim = torch.zeros(F, dtype=torch.float32, device="cuda")
    for rnd in range(2):
        dSums.zero_(); dSsim.zero_()
        api.compute_device(0, torch.cuda.current_stream().cuda_stream, W, H, 0, H, F, dA.data_ptr(), W, W * H, dB.data_ptr(), W, W * H,
                           None, W, W * H, dSums.data_ptr(), dSsim.data_ptr())
        torch.cuda.synchronize()
        for f in sorted({0, F // 2, F - 1}):
            o, tot, _ = oracle.oracle_ssim(a[f], b[f])
            assert abs(float(dSsim[f]) - float(o)) <= GLOBAL_TOL, (f, rnd, float(dSsim[f]), float(o))
            assert abs(float(dSums[f]) - tot) <= 2e-6 * W * H, (f, rnd)
        assert float(dSums[F // 2]) == float(W * H)
        if F > 1:
            assert float(dSsim[0]) < -0.5


def test_tuning_knobs_do_not_change_results():
    """ssim_cuda_set_tuning: any partition of the work (fewer warp pairs per SM, any minimum share) gives the same map values up
    to the per-piece centring (different pieces, same math)"""
    lib = api.cuda_lib()
    try:
        for (W, H, F) in ((640, 360, 1), (1920, 1080, 3)):
            a = np.stack([synth_pair(W, H, 12 + f)[0] for f in range(F)])
            b = np.stack([synth_pair(W, H, 12 + f)[1] for f in range(F)])
            want = [oracle.oracle_ssim(a[f], b[f], want_map=True) for f in range(F)]
            dA, dB = _dev(a), _dev(b)
            dMap = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
            dSsim = torch.empty(F, dtype=torch.float32, device="cuda")
            for wave, rows in ((1, 0), (3, 0), (8, 200), (5, 5000), (2, 1), (0, 0)):
                lib.ssim_cuda_set_tuning(wave, rows)
                dMap.zero_(); dSsim.zero_()
                api.compute_device(0, None, W, H, 0, H, F, dA.data_ptr(), W, W * H, dB.data_ptr(), W, W * H, dMap.data_ptr(), W, W * H, None, dSsim.data_ptr())
                torch.cuda.synchronize()
                for f in range(F):
                    assert abs(float(dSsim[f]) - float(want[f][0])) <= GLOBAL_TOL, (W, H, F, wave, rows, f)
                    assert np.abs(dMap[f].cpu().numpy() - want[f][2]).max() <= PIXEL_TOL, (W, H, F, wave, rows, f)
    finally:
        lib.ssim_cuda_set_tuning(0, 0)


def test_current_device_is_left_alone():
    """entry points make their device current only for the duration of the call"""
    if api.cuda_lib().ssim_cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(1)
    try:
        a, b = synth_pair(300, 100, 1)
        api.compute_ssim(a, b)                                  # SSIM_CUDA_DEVICE unset -> device 0
        assert torch.cuda.current_device() == 1
        x = torch.ones(8, device="cuda")
        assert x.device.index == 1
    finally:
        torch.cuda.set_device(0)


def test_compute_strips_entry_point():
    """ssim_cuda_compute_strips(): host image split across the GPUs of this process (+ NCCL when there are >= 2)"""
    n = api.cuda_lib().ssim_cuda_device_count()
    a, b = synth_pair(500, 333, 8)
    o, _, om = oracle.oracle_ssim(a, b, want_map=True)
    for devices in ([0], list(range(min(n, 2))), list(range(n))):
        s, m = api.compute_strips(devices, a, b, want_map=True)
        assert abs(float(s) - float(o)) <= GLOBAL_TOL, devices
        assert np.abs(m - om).max() <= PIXEL_TOL
    out = C.c_float()
    devs = (C.c_int * 2)(0, 0)
    assert api.cuda_lib().ssim_cuda_compute_strips(2, devs, 500, 333, a.ctypes.data, 1, 500, b.ctypes.data, 1, 500, None, 0, 0, C.byref(out)) == 22


def test_device_pointers_through_the_reference_api():
    """rmgr_ssim_compute_ssim() accepts device pointers for images and map (detected per pointer)"""
    W, H = 320, 90
    a, b = synth_pair(W, H, 2)
    o, _, om = oracle.oracle_ssim(a, b, want_map=True)
    dA, dB = _dev(a), _dev(b)
    dMap = torch.zeros((H, W), dtype=torch.float32, device="cuda")
    p = api.Params()
    p.width, p.height = W, H
    p.imgA.topLeft, p.imgA.step, p.imgA.stride = dA.data_ptr(), 1, W
    p.imgB.topLeft, p.imgB.step, p.imgB.stride = dB.data_ptr(), 1, W
    p.ssimMap, p.ssimStep, p.ssimStride = dMap.data_ptr(), 1, W
    out = C.c_float()
    assert api.rmgr_lib().rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 0
    assert abs(out.value - float(o)) <= GLOBAL_TOL and np.abs(dMap.cpu().numpy() - om).max() <= PIXEL_TOL
    # interleaved device image (step 3) + misaligned base goes through the pack kernel
    rgb = np.zeros((H, W, 3), np.uint8)
    rgb[..., 1] = a
    rgb2 = np.zeros((H, W, 3), np.uint8)
    rgb2[..., 1] = b
    dR, dR2 = _dev(rgb), _dev(rgb2)
    p.imgA.topLeft, p.imgA.step, p.imgA.stride = dR.data_ptr() + 1, 3, 3 * W
    p.imgB.topLeft, p.imgB.step, p.imgB.stride = dR2.data_ptr() + 1, 3, 3 * W
    p.ssimMap = None
    assert api.rmgr_lib().rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 0
    assert abs(out.value - float(o)) <= GLOBAL_TOL
    # interleaved DEVICE map (ssimStep = 3): written in place by the fused kernel, the other two channels stay untouched
    dM3 = torch.full((H, W, 3), -7.0, dtype=torch.float32, device="cuda")
    p.ssimMap, p.ssimStep, p.ssimStride = dM3.data_ptr() + 4, 3, 3 * W
    assert api.rmgr_lib().rmgr_ssim_compute_ssim(C.byref(out), C.byref(p), None) == 0
    got = dM3.cpu().numpy()
    assert np.abs(got[..., 1] - om).max() <= PIXEL_TOL and (got[..., 0] == -7.0).all() and (got[..., 2] == -7.0).all()


def test_all_channels_in_one_pass(bbb360):
    """interleaved RGB, all channels from one upload == the per-channel calls (step = 3) == oracle"""
    a, b = np.ascontiguousarray(bbb360["jpg50"]), np.ascontiguousarray(bbb360["png"])       # (80, 640, 3)
    s, m = api.compute_channels(a, b, want_map=True)
    for ch in range(3):
        o, _, om = oracle.oracle_ssim(a, b, want_map=True, step_a=3, step_b=3, stride_a=640 * 3, stride_b=640 * 3, width=640, height=80, a_off=ch, b_off=ch)
        one, m1 = api.compute_ssim(a, b, want_map=True, width=640, height=80, step_a=3, step_b=3, stride_a=640 * 3, stride_b=640 * 3, a_off=ch, b_off=ch)
        assert abs(float(s[ch]) - float(o)) <= GLOBAL_TOL and np.abs(m[..., ch] - om).max() <= PIXEL_TOL
        # (same math, different work decomposition => per-item centring pixels differ: equal to rounding, not bitwise)
        assert abs(float(s[ch]) - float(one)) <= 2e-7 and np.abs(m[..., ch] - m1).max() <= 3e-4
    s2, _ = api.compute_channels(a, b)
    assert np.array_equal(s, s2)
    g = np.ascontiguousarray(a[..., :1])
    s1, _ = api.compute_channels(g, np.ascontiguousarray(b[..., :1]))
    assert abs(float(s1[0]) - float(s[0])) <= 2e-7


def test_device_argument_errors():
    lib = api.cuda_lib()
    t = torch.zeros(64 * 64, dtype=torch.uint8, device="cuda")
    s = torch.zeros(1, dtype=torch.float64, device="cuda")
    f = lib.ssim_cuda_compute_device
    assert f(0, None, 64, 64, 0, 64, 1, t.data_ptr(), 64, 0, t.data_ptr(), 64, 0, None, 0, 0, None, None) == 22      # no output
    assert f(0, None, 64, 64, 0, 65, 1, t.data_ptr(), 64, 0, t.data_ptr(), 64, 0, None, 0, 0, s.data_ptr(), None) == 22  # rows out of range
    assert f(0, None, 60, 64, 0, 64, 1, t.data_ptr(), 60, 0, t.data_ptr(), 60, 0, None, 0, 0, s.data_ptr(), None) == 22  # pitch % 16
    assert f(0, None, 64, 64, 0, 64, 1, t.data_ptr() + 4, 64, 0, t.data_ptr(), 64, 0, None, 0, 0, s.data_ptr(), None) == 22  # base % 16
    assert f(99, None, 64, 64, 0, 64, 1, t.data_ptr(), 64, 0, t.data_ptr(), 64, 0, None, 0, 0, s.data_ptr(), None) == 22  # bad device


def test_concurrent_calls_from_many_threads():
    """the reference is re-entrant (SURVEY 8b 'Threading'); the GPU path serialises per device internally"""
    import threading
    pairs = [synth_pair(300 + 17 * i, 120 + 5 * i, i) for i in range(6)]
    want = [oracle.oracle_ssim(a, b, want_map=True) for a, b in pairs]
    errors = []

    def worker(k):
        try:
            for _ in range(5):
                a, b = pairs[k]
                s, m = api.compute_ssim(a, b, want_map=True)
                if abs(float(s) - float(want[k][0])) > GLOBAL_TOL or np.abs(m - want[k][2]).max() > PIXEL_TOL:
                    errors.append((k, float(s), float(want[k][0])))
        except Exception as e:            # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    # two streams on the device API at the same time (separate partial-sum workspaces per stream)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for st, (a, b) in zip((s1, s2), pairs[:2]):
        h, w = a.shape
        pw = (w + 15) // 16 * 16
        da = torch.zeros((h, pw), dtype=torch.uint8, device="cuda"); db = torch.zeros_like(da)
        da[:, :w] = _dev(a); db[:, :w] = _dev(b)
        ds = torch.empty(1, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        api.compute_device(0, st.cuda_stream, w, h, 0, h, 1, da.data_ptr(), pw, 0, db.data_ptr(), pw, 0, None, 0, 0, None, ds.data_ptr())
        outs.append((ds, da, db))
    torch.cuda.synchronize()
    for (ds, _, _), wv in zip(outs, want[:2]):
        assert abs(float(ds.item()) - float(wv[0])) <= GLOBAL_TOL


def test_full_size_known_answers(golden):
    """BASELINE.json configs[3]/[4] at their FULL sizes against values the unmodified reference produced here
    (tests/golden/make_golden.py): the 16384x16384 pair as one image and as 8 row strips with halos, and frames
    0 / 1 / 4095 of the 1080p sweep inside one batched launch.  Inputs come from the device-side synthetic recipe,
    whose equality with the host recipe is checked in test_batch_of_frames_matches_per_frame_oracle."""
    import torch
    from ssim_b200 import parallel
    st = torch.cuda.current_stream().cuda_stream
    W = H = 16384
    a = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    api.synth_fill(0, st, a.data_ptr(), W, b.data_ptr(), W, W, H, 0, 0)
    val = torch.empty(1, dtype=torch.float32, device="cuda")
    sums = torch.empty(1, dtype=torch.float64, device="cuda")
    api.compute_device(0, st, W, H, 0, H, 1, a.data_ptr(), W, 0, b.data_ptr(), W, 0, None, 0, 0, sums.data_ptr(), val.data_ptr())
    torch.cuda.synchronize()
    want = golden["synthetic"]["16384x16384_f0"]["ref_f64_auto"]
    assert abs(float(val.item()) - want) <= GLOBAL_TOL
    # 8 strips with 5 halo rows on interior edges: the partial sums add up to the single-image sum
    total = 0.0
    part = torch.empty(1, dtype=torch.float64, device="cuda")
    for r in range(8):
        s0, s1, oy, orows = parallel.strip_bounds(H, 8, r)
        api.compute_device(0, st, W, s1 - s0, oy, orows, 1, a[s0:s1].data_ptr(), W, 0, b[s0:s1].data_ptr(), W, 0, None, 0, 0, part.data_ptr(), None)
        torch.cuda.synchronize()
        total += float(part.item())
    assert abs(parallel.mean_from_partials(total, W, H) - want) <= GLOBAL_TOL
    assert abs(total - float(sums.item())) <= 2e-7 * W * H
    del a, b
    # 1080p frames 0, 1, 4095 in one launch
    w, h = 1920, 1080
    frames = [0, 1, 4095]
    fa = torch.empty((3, h, w), dtype=torch.uint8, device="cuda")
    fb = torch.empty_like(fa)
    for i, f in enumerate(frames):
        api.synth_fill(0, st, fa[i].data_ptr(), w, fb[i].data_ptr(), w, w, h, 0, f)
    fv = torch.empty(3, dtype=torch.float32, device="cuda")
    api.compute_device(0, st, w, h, 0, h, 3, fa.data_ptr(), w, w * h, fb.data_ptr(), w, w * h, None, 0, 0, None, fv.data_ptr())
    torch.cuda.synchronize()
    for i, f in enumerate(frames):
        assert abs(float(fv[i].item()) - golden["synthetic"]["1920x1080_f%d" % f]["ref_f64_auto"]) <= GLOBAL_TOL


def test_calls_can_be_captured_into_a_cuda_graph():
    """ssim_cuda_compute_device() only enqueues work on the caller's stream (one kernel launch; workspace and descriptors are
    in place after the first call on that stream), so a sequence of calls can be captured into a CUDA graph and replayed."""
    W, H, n = 640, 360, 4
    pairs = [synth_pair(W, H, f) for f in range(n)]
    dA = [_dev(p[0]) for p in pairs]
    dB = [_dev(p[1]) for p in pairs]
    val = torch.zeros(n, dtype=torch.float32, device="cuda")
    maps = torch.zeros((n, H, W), dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())

    def enqueue():
        for i in range(n):
            api.compute_device(0, s.cuda_stream, W, H, 0, H, 1, dA[i].data_ptr(), W, 0, dB[i].data_ptr(), W, 0,
                               maps[i].data_ptr(), W, 0, None, val[i:].data_ptr())

    enqueue()                                             # first use of the stream: allocates its workspace
    torch.cuda.synchronize()
    want = val.clone(); want_maps = maps.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        enqueue()
    for _ in range(2):
        val.zero_(); maps.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(val, want) and torch.equal(maps, want_maps)
    for i in range(n):
        assert abs(float(val[i]) - float(oracle.oracle_ssim(*pairs[i])[0])) <= GLOBAL_TOL


def test_two_planes_sharing_a_descriptor_cache_entry():
    """The tensor-map descriptors of a call come from a direct-mapped per-thread cache; A and B of one call may fall into the
    same entry (1 call in 512 with unrelated addresses).  Build such a pair on purpose: the second look-up must not change what
    the first one returned (a round-2 bug handed out pointers into the cache: the call then compared B with B, SSIM = 1)."""
    lib = api.cuda_lib()
    W, H = 208, 77
    plane = W * H                                         # 16016 bytes: a multiple of 16
    n = 4096
    pool = torch.empty(n * plane, dtype=torch.uint8, device="cuda")
    seen = {}
    pair = None
    for i in range(n):
        e = lib.ssim_cuda_debug_map_cache_entry(pool.data_ptr() + i * plane, W, H, 1, W, 1)
        if e in seen:
            pair = (seen[e], i)
            break
        seen[e] = i
    assert pair is not None
    a, b = synth_pair(W, H, 3)
    pool[pair[0] * plane:(pair[0] + 1) * plane] = _dev(a).flatten()
    pool[pair[1] * plane:(pair[1] + 1) * plane] = _dev(b).flatten()
    val = torch.zeros(1, dtype=torch.float32, device="cuda")
    want = float(oracle.oracle_ssim(a, b)[0])
    for _ in range(2):                                    # cold entry, then whatever the first call left in it
        api.compute_device(0, torch.cuda.current_stream().cuda_stream, W, H, 0, H, 1, pool.data_ptr() + pair[0] * plane, W, 0,
                           pool.data_ptr() + pair[1] * plane, W, 0, None, 0, 0, None, val.data_ptr())
        torch.cuda.synchronize()
        assert abs(float(val.item()) - want) <= GLOBAL_TOL, (float(val.item()), want)


def test_full_size_properties_4k():
    """Size-independent properties at the headline size (3840x2160, BASELINE.json configs[2]): SSIM(a, b) == SSIM(b, a) bit for
    bit, map included (the kernel's formula is symmetric in its two inputs); the global sum is the sum of the stored map; a
    frame scores the same alone and as any member of a batch (other partition, other centring pixels: to 2e-7)."""
    st = torch.cuda.current_stream().cuda_stream
    W, H, F = 3840, 2160, 3
    a = torch.empty((F, H, W), dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    for f in range(F):
        api.synth_fill(0, st, a[f].data_ptr(), W, b[f].data_ptr(), W, W, H, 0, 5)        # three copies of frame 5
    m1 = torch.empty((H, W), dtype=torch.float32, device="cuda")
    m2 = torch.empty_like(m1)
    s1 = torch.empty(1, dtype=torch.float64, device="cuda"); s2 = torch.empty_like(s1)
    v1 = torch.empty(1, dtype=torch.float32, device="cuda"); v2 = torch.empty_like(v1)
    api.compute_device(0, st, W, H, 0, H, 1, a.data_ptr(), W, 0, b.data_ptr(), W, 0, m1.data_ptr(), W, 0, s1.data_ptr(), v1.data_ptr())
    api.compute_device(0, st, W, H, 0, H, 1, b.data_ptr(), W, 0, a.data_ptr(), W, 0, m2.data_ptr(), W, 0, s2.data_ptr(), v2.data_ptr())
    torch.cuda.synchronize()
    assert float(v1.item()) == float(v2.item()) and float(s1.item()) == float(s2.item())
    assert torch.equal(m1, m2)
    assert abs(float(m1.double().sum().item()) - float(s1.item())) <= 1e-7 * W * H
    vb = torch.empty(F, dtype=torch.float32, device="cuda")
    sb = torch.empty(F, dtype=torch.float64, device="cuda")
    api.compute_device(0, st, W, H, 0, H, F, a.data_ptr(), W, W * H, b.data_ptr(), W, W * H, None, 0, 0, sb.data_ptr(), vb.data_ptr())
    torch.cuda.synchronize()
    for f in range(F):
        assert abs(float(vb[f].item()) - float(v1.item())) <= 2e-7, f
        assert abs(float(sb[f].item()) - float(s1.item())) <= 2e-7 * W * H, f


def test_strip_sums_exchanged_through_peer_memory():
    """ssim_cuda_compute_strip_allreduce(): the kernel of every rank stores its strip sum into every peer's exchange
    buffer and adds up what lands in its own.  One process drives all visible GPUs (ranks = devices; with a single GPU the
    ranks share device 0, which still exercises the slot / epoch protocol); every rank must end with the bits of the full-image
    result, twice in a row (epoch parity), and a missing peer must time out instead of hanging."""
    import torch
    n_dev = torch.cuda.device_count()
    world = 4
    W, H = 640, 333
    a, b = synth_pair(W, H, 2)
    o, osum, _ = oracle.oracle_ssim(a, b)
    devs = [r % n_dev for r in range(world)]
    lib = api.cuda_lib()
    bufs = [api.exchange_create(d)[0] for d in devs]
    for d in set(devs):
        for e in set(devs):
            assert lib.ssim_cuda_exchange_enable_peer(d, e) == 0
    state = []
    for r in range(world):
        s0, s1, oy, orows = parallel.strip_bounds(H, world, r)
        dev = torch.device("cuda", devs[r])
        da = torch.from_numpy(np.ascontiguousarray(a[s0:s1])).to(dev); db = torch.from_numpy(np.ascontiguousarray(b[s0:s1])).to(dev)
        state.append((da, db, s1 - s0, oy, orows, torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float32, device=dev),
                      torch.full((1,), -1, dtype=torch.int32, device=dev), torch.cuda.Stream(device=dev)))
    for epoch in (1, 2, 3):
        for r in range(world):
            da, db, rows, oy, orows, dsum, dval, dst, st = state[r]
            api.compute_strip_allreduce(devs[r], st.cuda_stream, W, rows, oy, orows, H, da.data_ptr(), W, db.data_ptr(), W, None, 0,
                                        bufs, r, epoch, dsum.data_ptr(), dval.data_ptr(), dst.data_ptr())
        for d in set(devs):
            torch.cuda.synchronize(d)
        vals = [float(s[6].item()) for s in state]
        sums = [float(s[5].item()) for s in state]
        assert all(int(s[7].item()) == 0 for s in state)
        assert len(set(sums)) == 1 and len(set(vals)) == 1                      # identical bits on every rank
        assert abs(vals[0] - float(o)) <= GLOBAL_TOL and abs(sums[0] - osum) <= 2e-7 * W * H
    # a rank whose peers never arrive gives up after the time-out (status 1, NaN), it does not hang
    import os
    da, db, rows, oy, orows, dsum, dval, dst, st = state[0]
    lonely = [api.exchange_create(devs[0])[0] for _ in range(2)]
    os.environ["SSIM_CUDA_EXCHANGE_TIMEOUT_MS"] = "2000"
    api.compute_strip_allreduce(devs[0], st.cuda_stream, W, rows, oy, orows, H, da.data_ptr(), W, db.data_ptr(), W, None, 0,
                                lonely, 0, 1, dsum.data_ptr(), dval.data_ptr(), dst.data_ptr())
    torch.cuda.synchronize(devs[0])
    assert int(dst.item()) == 1 and np.isnan(float(dsum.item()))
    for bptr, d in zip(bufs + lonely, devs + [devs[0]] * 2):
        assert lib.ssim_cuda_exchange_destroy(d, bptr) == 0
