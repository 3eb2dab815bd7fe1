import sys; sys.path.insert(0,'/root/repo')
import torch
from ssim_b200 import api
st = torch.cuda.current_stream().cuda_stream
W, H = 3840, 2160
a = torch.empty((H, W), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
api.synth_fill(0, st, a.data_ptr(), W, b.data_ptr(), W, W, H, 0, 5)
m1 = torch.empty((H, W), dtype=torch.float32, device="cuda"); m2 = torch.empty_like(m1)
s1 = torch.empty(1, dtype=torch.float64, device="cuda"); s2 = torch.empty_like(s1)
v1 = torch.empty(1, dtype=torch.float32, device="cuda"); v2 = torch.empty_like(v1)
bad = 0
ref = None
for it in range(300):
    m1.fill_(-7.0); m2.fill_(-7.0)
    api.compute_device(0, st, W, H, 0, H, 1, a.data_ptr(), W, 0, b.data_ptr(), W, 0, m1.data_ptr(), W, 0, s1.data_ptr(), v1.data_ptr())
    api.compute_device(0, st, W, H, 0, H, 1, b.data_ptr(), W, 0, a.data_ptr(), W, 0, m2.data_ptr(), W, 0, s2.data_ptr(), v2.data_ptr())
    torch.cuda.synchronize()
    if ref is None: ref = (m1.clone(), float(s1.item()))
    e1 = not torch.equal(m1, m2); e2 = float(s1.item()) != float(s2.item()); e3 = not torch.equal(m1, ref[0]); e4 = float(s1.item()) != ref[1]
    if e1 or e2 or e3 or e4:
        bad += 1
        d = (m1 != m2).nonzero()
        d3 = (m1 != ref[0]).nonzero()
        print(it, 'map sym diff', e1, d.shape[0], d[:3].tolist(), 'sum diff', e2, float(s1.item()), float(s2.item()), 'vs first', e3, d3.shape[0], d3[:3].tolist(), e4, flush=True)
print('iterations 300 bad', bad)
