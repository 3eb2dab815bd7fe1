"""Dev A/B: times the fused launch of several builds of libssim_cuda.so (build/var_<name>/, made with -DSSIM_VAR_* flags)
on the headline batch and on single images, each build in its own process, and prints a hash of the map so that variants
that must be bit-identical can be checked.   python tools/dev/variant_ab.py base f16 ..."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r'''
import sys, hashlib
sys.path.insert(0, %(root)r)
import torch
from ssim_b200 import api
api.LIB_DIR = %(libdir)r
lib = api.cuda_lib()
st = torch.cuda.current_stream(); sh = st.cuda_stream
def run(W, H, F, with_map, n, reps=3):
    a = torch.empty((F, H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
    for f in range(F): api.synth_fill(0, sh, a[f].data_ptr(), W, b[f].data_ptr(), W, W, H, 0, f)
    m = torch.zeros((F, H, W), dtype=torch.float32, device='cuda') if with_map else None
    sums = torch.empty(F, dtype=torch.float64, device='cuda')
    def fn(): api.compute_device(0, sh, W, H, 0, H, F, a.data_ptr(), W, W * H, b.data_ptr(), W, W * H, m.data_ptr() if with_map else None, W, W * H, sums.data_ptr(), None)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(n): fn()
        e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    h = hashlib.sha1(m[0].cpu().numpy().tobytes()).hexdigest()[:10] if with_map else "-"
    return best, sums[0].item(), h
for name, W, H, F, mp, n in [("64x4K+map", 3840, 2160, 64, True, 10), ("16x4K+map", 3840, 2160, 16, True, 20), ("4K+map", 3840, 2160, 1, True, 50),
                             ("1080p nomap", 1920, 1080, 1, False, 50), ("1296x717+map", 1296, 717, 3, True, 50)]:
    us, s0, h = run(W, H, F, mp, n)
    print("%%-11s %%-14s %%9.2f us  %%9.0f Mpix/s  sum0 %%.9f  map %%s" %% (%(name)r, name, us, W * H * F / us, s0, h), flush=True)
'''

for name in sys.argv[1:]:
    # "variant@ns": the same build with SSIM_CUDA_BACKOFF_NS=ns
    base, _, backoff = name.partition("@")
    libdir = os.path.join(ROOT, "build", "var_" + base)
    env = dict(os.environ)
    if backoff:
        env["SSIM_CUDA_BACKOFF_NS"] = backoff
    # the engine is loaded from the variant directory; librmgr-ssim.so is not needed here
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "libdir": libdir, "name": name}], check=False, env=env)
