"""Dev: N calls of ssim_cuda_compute_device captured into ONE CUDA graph and replayed (queued small calls without any host work)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from ssim_b200 import api


def run(W, H, with_map, n=20):
    a = torch.empty((H, W), dtype=torch.uint8, device='cuda'); b = torch.empty_like(a)
    m = torch.empty((H, W), dtype=torch.float32, device='cuda') if with_map else None
    val = torch.empty(n, dtype=torch.float32, device='cuda')
    st0 = torch.cuda.current_stream()
    api.synth_fill(0, st0.cuda_stream, a.data_ptr(), W, b.data_ptr(), W, W, H, 0, 0)
    s = torch.cuda.Stream()
    s.wait_stream(st0)
    with torch.cuda.stream(s):
        for i in range(3):      # warm: workspace of this stream, descriptors
            api.compute_device(0, s.cuda_stream, W, H, 0, H, 1, a.data_ptr(), W, 0, b.data_ptr(), W, 0, m.data_ptr() if with_map else None, W, 0, None, val[i:].data_ptr())
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(n):
            api.compute_device(0, s.cuda_stream, W, H, 0, H, 1, a.data_ptr(), W, 0, b.data_ptr(), W, 0, m.data_ptr() if with_map else None, W, 0, None, val[i:].data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record(); 
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) * 1e3 / (10 * n)
    # the same calls queued directly
    with torch.cuda.stream(s):
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(s)
        for _ in range(10):
            for i in range(n):
                api.compute_device(0, s.cuda_stream, W, H, 0, H, 1, a.data_ptr(), W, 0, b.data_ptr(), W, 0, m.data_ptr() if with_map else None, W, 0, None, val[i:].data_ptr())
        f1.record(s)
    torch.cuda.synchronize()
    direct = f0.elapsed_time(f1) * 1e3 / (10 * n)
    print("%dx%d map=%d: graph replay %.2f us per call, direct queueing %.2f us per call, ssim %.6f" % (W, H, with_map, per, direct, float(val[n - 1])))


run(1920, 1080, False)
run(3840, 2160, True)
run(256, 256, True)
