"""Dev sweep: 64 x 4K pairs with map, device-timed, vs rows per work item (segment).  SSIM_CUDA_BACKOFF_NS is read once per process."""
import sys; sys.path.insert(0, '/root/repo')
import os, statistics, torch
from ssim_b200 import api
lib = api.cuda_lib()
st = torch.cuda.current_stream(); sh = st.cuda_stream
W, H, F = 3840, 2160, 64
a = torch.empty((F, H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
m = torch.empty((F, H, W), dtype=torch.float32, device='cuda')
sums = torch.empty(F, dtype=torch.float64, device='cuda')
for f in range(F): api.synth_fill(0, sh, a[f].data_ptr(), W, b[f].data_ptr(), W, W, H, 0, f)
segs = [int(x) for x in sys.argv[1:]] or [0]
for seg in segs:
    lib.ssim_cuda_set_segment_rows(seg)
    ts = []
    for i in range(13):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        api.compute_device(0, sh, W, H, 0, H, F, a.data_ptr(), W, W * H, b.data_ptr(), W, W * H, m.data_ptr(), W, W * H, sums.data_ptr(), None)
        e1.record(st); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    t = statistics.median(ts)
    print("backoff", os.environ.get("SSIM_CUDA_BACKOFF_NS", "default"), "segRows", seg, "ms %.4f" % t, "Mpix/s %.0f" % (W * H * F / t / 1e3), flush=True)
lib.ssim_cuda_set_segment_rows(0)
