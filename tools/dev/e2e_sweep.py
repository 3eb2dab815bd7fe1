import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from ssim_b200 import api
W, H = 3840, 2160
hA = torch.empty((4, H, W), dtype=torch.uint8).pin_memory(); hB = torch.empty_like(hA).pin_memory(); hM = torch.empty((4, H, W), dtype=torch.float32).pin_memory()
d = torch.empty((H, W), dtype=torch.uint8, device="cuda"); e = torch.empty_like(d)
for f in range(4):
    api.synth_fill(0, None, d.data_ptr(), W, e.data_ptr(), W, W, H, 0, f); torch.cuda.synchronize(); hA[f].copy_(d); hB[f].copy_(e)
nA, nB, nM = hA.numpy(), hB.numpy(), hM.numpy()
def run(n=40, want_map=True):
    for k in range(4): api.compute_ssim(nA[k], nB[k], ssim_map=nM[k] if want_map else None)
    t0 = time.perf_counter()
    for i in range(n): s, _ = api.compute_ssim(nA[i & 3], nB[i & 3], ssim_map=nM[i & 3] if want_map else None)
    dt = (time.perf_counter() - t0) / n
    return dt * 1e3, W * H / dt / 1e6, float(s)
print("chunkKB", os.environ.get("SSIM_CUDA_CHUNK_KB"), "nopipe", os.environ.get("SSIM_CUDA_NO_PIPELINE"), "map: %.3f ms %.0f Mpix/s %.6f" % run(), " nomap: %.3f ms %.0f Mpix/s" % run(want_map=False)[:2])
