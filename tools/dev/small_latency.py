"""End-to-end latency of one blocking rmgr_ssim_compute_ssim call (host buffers) for small images, next to the reference."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from ssim_b200 import api
from ssim_b200.synth import synth_pair
import oracle
def med(fn, n=200):
    for _ in range(10): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort(); return ts[len(ts) // 2] * 1e6
for (w, h) in [(64, 64), (256, 256), (640, 360), (1280, 720), (1920, 1080)]:
    a, b = synth_pair(w, h, 1)
    m = np.empty((h, w), np.float32)
    g_map = med(lambda: api.compute_ssim(a, b, want_map=True, ssim_map=m))
    g_nomap = med(lambda: api.compute_ssim(a, b, want_map=False))
    r_map = med(lambda: oracle.ref_ssim("f32", a, b, want_map=True, openmp=True), 50)
    r_nomap = med(lambda: oracle.ref_ssim("f32", a, b, want_map=False, openmp=True), 50)
    r1 = med(lambda: oracle.ref_ssim("f32", a, b, want_map=False, openmp=False), 20)
    print("%4dx%-4d  gpu map %7.1f us  no-map %7.1f us | reference openmp map %7.1f  no-map %7.1f  serial no-map %8.1f" % (w, h, g_map, g_nomap, r_map, r_nomap, r1), flush=True)
