"""Determinism stress: repeats device-resident computations and counts runs whose per-frame sums differ bitwise from the first."""
import sys; sys.path.insert(0, '/root/repo')
import torch
from ssim_b200 import api
st = torch.cuda.current_stream(); sh = st.cuda_stream
def stress(W, H, F, with_map, reps):
    a = torch.empty((F, H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
    m = torch.empty((F, H, W), dtype=torch.float32, device='cuda') if with_map else None
    sums = torch.empty(F, dtype=torch.float64, device='cuda')
    for f in range(F): api.synth_fill(0, sh, a[f].data_ptr(), W, b[f].data_ptr(), W, W, H, 0, f)
    ref = None; refm = None; bad = 0; badm = 0
    for i in range(reps):
        api.compute_device(0, sh, W, H, 0, H, F, a.data_ptr(), W, W * H, b.data_ptr(), W, W * H, m.data_ptr() if with_map else None, W, W * H, sums.data_ptr(), None)
        torch.cuda.synchronize()
        s = sums.clone()
        if ref is None: ref = s; refm = m.clone() if with_map else None
        else:
            if not torch.equal(s, ref): bad += 1
            if with_map and not torch.equal(m, refm): badm += 1
    print("%dx%d x%d map=%d: %d/%d runs differ (sums), %d (map)" % (W, H, F, with_map, bad, reps - 1, badm), flush=True)
    return bad + badm
tot = 0
tot += stress(256, 256, 1, True, 400)
tot += stress(1920, 1080, 1, False, 300)
tot += stress(1920, 1080, 1, True, 300)
tot += stress(3840, 2160, 1, True, 200)
tot += stress(336, 141, 7, True, 300)
tot += stress(3840, 2160, 16, True, 40)
print("TOTAL", tot)
