"""Dev: 64 x 4K batch and one 4K pair, device-timed, with whichever ssim_b200 package is first on sys.path (argv[1] = repo root)."""
import sys
sys.path.insert(0, sys.argv[1])
import torch
from ssim_b200 import api
print("library:", api.LIB_DIR)
st = torch.cuda.current_stream(); sh = st.cuda_stream
def run(F, reps):
    W, H = 3840, 2160
    a = torch.empty((F, H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
    m = torch.empty((F, H, W), dtype=torch.float32, device='cuda')
    sums = torch.empty(F, dtype=torch.float64, device='cuda'); val = torch.empty(F, dtype=torch.float32, device='cuda')
    for f in range(F):
        api.synth_fill(0, sh, a[f].data_ptr(), W, b[f].data_ptr(), W, W, H, 0, f)
    fn = lambda: api.compute_device(0, sh, W, H, 0, H, F, a.data_ptr(), W, W * H, b.data_ptr(), W, W * H, m.data_ptr(), W, W * H, sums.data_ptr(), val.data_ptr())
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print("%2d x 4K: %8.1f us per call = %8.0f Mpix/s  ssim %.6f" % (F, us, F * W * H / us, float(val[0])))
run(16, 20); run(64, 10); run(16, 20); run(128, 5)
