"""Dev: per-warp-pair start/finish times of one launch of the persistent kernel (load balance of the static partition)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from ssim_b200 import api
lib = api.cuda_lib()
st = torch.cuda.current_stream(); sh = st.cuda_stream
def run(W, H, F, with_map=True):
    a = torch.empty((F, H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
    m = torch.empty((F, H, W), dtype=torch.float32, device='cuda') if with_map else None
    sums = torch.empty(F, dtype=torch.float64, device='cuda')
    for f in range(F):
        api.synth_fill(0, sh, a[f].data_ptr(), W, b[f].data_ptr(), W, W, H, 0, f)
    times = torch.zeros(32 * 1184, dtype=torch.int64, device='cuda')
    def call():
        api.compute_device(0, sh, W, H, 0, H, F, a.data_ptr(), W, W * H, b.data_ptr(), W, W * H, m.data_ptr() if with_map else None, W, W * H, sums.data_ptr(), None)
    for _ in range(3): call()
    torch.cuda.synchronize()
    lib.ssim_cuda_debug_slot_times(times.data_ptr())
    call(); torch.cuda.synchronize()
    lib.ssim_cuda_debug_slot_times(None)
    full = times.cpu().numpy().reshape(-1, 32).astype(np.float64)
    full = full[full[:, 1] > 0]
    t = full[:, :2]
    t0 = t[:, 0].min()
    start = (t[:, 0] - t0) / 1e3; end = (t[:, 1] - t0) / 1e3
    q = lambda x: np.percentile(x, [0, 5, 25, 50, 75, 95, 100]).round(1)
    print("%dx%d x%d map=%d: %d slots; start us %s; finish us %s" % (W, H, F, with_map, len(t), q(start), q(end)))
    # per SMSP position: slot s -> CTA s//4 (SM unknown), pair s%4
    PAIRS = 8
    for pr in range(PAIRS):
        print("   pair %d finish median %.1f" % (pr, np.median(end[pr::PAIRS])), end="")
    print()
    # first half of the CTAs (launched first) vs second half
    cta = np.arange(len(t)) // PAIRS
    per_cta = [float(np.median(end[cta == c])) for c in range(int(cta.max()) + 1)]
    print("   finish per CTA (median of its pairs):", " ".join("%.0f" % v for v in per_cta))
    order = np.argsort(-end)[:12]
    print("   slowest slots (slot: start, finish us):", " ".join("%d: %.1f-%.1f" % (i, start[i], end[i]) for i in order))

if len(sys.argv) > 3:
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), len(sys.argv) <= 4)
else:
    run(3840, 2160, 1)
    run(3840, 2160, 16)
    run(3840, 2160, 64)
    run(1920, 1080, 1, False)
