import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from ssim_b200.synth import synth_pair
frames = [synth_pair(3840, 2160, f) for f in range(2)]
for rep in range(3):
    mp, cores, done, dt, s = bench.time_reference(frames, steps=10, warmup=2)
    print("numpy frames: %.0f Mpix/s (%d calls, %.2f s) cores=%d" % (mp, done, dt, cores))
if len(sys.argv) > 1:
    import torch
    pa = [(torch.from_numpy(a).pin_memory().numpy(), torch.from_numpy(b).pin_memory().numpy()) for a, b in frames]
    for rep in range(3):
        mp, cores, done, dt, s = bench.time_reference(pa, steps=10, warmup=2)
        print("pinned frames after torch import: %.0f Mpix/s" % mp)
    mp, cores, done, dt, s = bench.time_reference(frames, steps=10, warmup=2)
    print("numpy frames after torch import: %.0f Mpix/s" % mp)
print(os.environ.get("OMP_NUM_THREADS"), os.cpu_count(), len(os.sched_getaffinity(0)))
