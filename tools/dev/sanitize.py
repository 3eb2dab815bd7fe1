"""Dev: a few small launches of every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from ssim_b200 import api
from ssim_b200.synth import synth_pair
lib = api.cuda_lib()
for (w, h) in ((333, 141), (68, 97), (640, 360), (1284, 31)):
    a, b = synth_pair(w, h, 2)
    print(w, h, api.compute_ssim(a, b, want_map=True)[0], api.compute_ssim(a, b)[0],
          api.compute_u16(a.astype(np.uint16) * 257, b.astype(np.uint16) * 257, want_map=True)[0])
a = np.stack([synth_pair(200, 90, 1)[0]] * 3, axis=-1).copy(); b = np.stack([synth_pair(200, 90, 1)[1]] * 3, axis=-1).copy()
print("channels", api.compute_channels(a, b, want_map=True)[0])
lib.ssim_cuda_set_tuning(4, 0)           # fewer warp pairs per SM than the default
a, b = synth_pair(640, 720, 5)
print("4 pairs per SM", api.compute_ssim(a, b, want_map=True)[0])
lib.ssim_cuda_set_tuning(0, 0)
# slots that walk through several pieces (many small frames in one launch) and the in-kernel per-frame reduction
import torch
F, W, H = 70, 208, 77
fa = torch.from_numpy(np.stack([synth_pair(W, H, f)[0] for f in range(F)])).cuda()
fb = torch.from_numpy(np.stack([synth_pair(W, H, f)[1] for f in range(F)])).cuda()
fm = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
fs = torch.empty(F, dtype=torch.float32, device="cuda")
for _ in range(2):
    api.compute_device(0, torch.cuda.current_stream().cuda_stream, W, H, 0, H, F, fa.data_ptr(), W, W * H, fb.data_ptr(), W, W * H,
                       fm.data_ptr(), W, W * H, None, fs.data_ptr())
torch.cuda.synchronize()
print("70 frames in one launch", float(fs[0]), float(fs[F - 1]))
