"""Dev driver for ncu: launches of the fused kernel on ONE pair (argv: width height launches [nomap])."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from ssim_b200 import api
W = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
N = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nomap = len(sys.argv) > 4
st = torch.cuda.current_stream(); sh = st.cuda_stream
a = torch.empty((H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
m = torch.empty((H, W), dtype=torch.float32, device='cuda')
sums = torch.empty(1, dtype=torch.float64, device='cuda')
api.synth_fill(0, sh, a.data_ptr(), W, b.data_ptr(), W, W, H, 0, 0)
for _ in range(N):
    api.compute_device(0, sh, W, H, 0, H, 1, a.data_ptr(), W, 0, b.data_ptr(), W, 0, None if nomap else m.data_ptr(), W, 0, sums.data_ptr(), None)
torch.cuda.synchronize()
print(float(sums[0].item()))
