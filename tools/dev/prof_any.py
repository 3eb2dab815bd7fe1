"""Dev driver for ncu with an arbitrary checkout: argv = repo root, frames, launches."""
import sys
sys.path.insert(0, sys.argv[1])
import torch
from ssim_b200 import api
F = int(sys.argv[2]); N = int(sys.argv[3])
W, H = 3840, 2160
st = torch.cuda.current_stream(); sh = st.cuda_stream
a = torch.empty((F, H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
m = torch.empty((F, H, W), dtype=torch.float32, device='cuda')
sums = torch.empty(F, dtype=torch.float64, device='cuda')
for f in range(F):
    api.synth_fill(0, sh, a[f].data_ptr(), W, b[f].data_ptr(), W, W, H, 0, f)
for _ in range(N):
    api.compute_device(0, sh, W, H, 0, H, F, a.data_ptr(), W, W * H, b.data_ptr(), W, W * H, m.data_ptr(), W, W * H, sums.data_ptr(), None)
torch.cuda.synchronize()
print(float(sums[0].item()))
