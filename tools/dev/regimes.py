"""Dev sweep: device time of the fused launch in the regimes VERDICT r1 names, vs the two tuning knobs.

  single 4K pair with map, single 1080p pair without map (latency: events around ONE call, and N calls queued back to back),
  one 16384 x 2058 strip with map (the per-GPU share at N = 8), 64 x 4K batch with map (throughput).
Usage: python tools/dev/regimes.py [quick]
"""
import statistics
import sys

sys.path.insert(0, '/root/repo')
import torch

from ssim_b200 import api

lib = api.cuda_lib()
st = torch.cuda.current_stream()
sh = st.cuda_stream


def planes(W, rows, F):
    a = torch.empty((F, rows, W), dtype=torch.uint8, device='cuda')
    b = torch.empty_like(a)
    for f in range(F):
        api.synth_fill(0, sh, a[f].data_ptr(), W, b[f].data_ptr(), W, W, rows, 0, f)
    return a, b


def one_call_us(fn, n=30):
    ts = []
    for i in range(n + 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        fn()
        e1.record(st)
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)


def queued_us(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def case(name, W, src_rows, oy, orows, F, with_map, knobs, queued_n=50):
    a, b = planes(W, src_rows, F)
    m = torch.empty((F, orows, W), dtype=torch.float32, device='cuda') if with_map else None
    sums = torch.empty(F, dtype=torch.float64, device='cuda')
    val = torch.empty(F, dtype=torch.float32, device='cuda')

    def fn():
        api.compute_device(0, sh, W, src_rows, oy, orows, F, a.data_ptr(), W, W * src_rows, b.data_ptr(), W, W * src_rows,
                           m.data_ptr() if with_map else None, W, W * orows, sums.data_ptr(), val.data_ptr())

    for pairs, min_rows in knobs:
        lib.ssim_cuda_set_tuning(pairs, min_rows)
        med, mn = one_call_us(fn)
        q = queued_us(fn, queued_n)
        px = W * orows * F
        print("%-28s pairs/SM %d minRows %3d : one call median %8.1f us (min %8.1f)  queued %8.1f us/call = %9.0f Mpix/s  ssim %.6f" %
              (name, pairs or 8, min_rows or 6, med, mn, q, px / q, float(val[0].item())), flush=True)
    lib.ssim_cuda_set_tuning(0, 0)


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
case("4K pair + map", 3840, 2160, 0, 2160, 1, True, [(0, 0)] if quick else [(0, 0), (0, 48), (0, 72)])
case("1080p pair no map", 1920, 1080, 0, 1080, 1, False, [(0, 0)] if quick else [(0, 0), (0, 12), (0, 16), (0, 32), (0, 48)])
case("1080p pair + map", 1920, 1080, 0, 1080, 1, True, [(0, 0)] if quick else [(0, 0), (0, 16), (0, 32), (0, 48)])
case("256x256 + map", 256, 256, 0, 256, 1, True, [(0, 0)] if quick else [(0, 0), (0, 12), (0, 16), (0, 48)])
case("16384x2058 strip + map", 16384, 2068, 5, 2058, 1, True, [(0, 0)] if quick else [(0, 0), (4, 0)], queued_n=20)
case("16 x 4K + map", 3840, 2160, 0, 2160, 16, True, [(0, 0)], queued_n=20)
case("64 x 4K + map", 3840, 2160, 0, 2160, 64, True, [(0, 0)] if quick else [(0, 0), (4, 0)], queued_n=10)
case("512 x 1080p + map", 1920, 1080, 0, 1080, 512, True, [(0, 0)], queued_n=5)
