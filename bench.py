#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: SSIM Mpix/s (device-timed), 3840x2160 8-bit pairs WITH per-pixel map.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's own FMA+OpenMP CPU path, oracle/_ref)

A "step" is one pass of the hot path over one batch: FRAMES synthetic 4K frame pairs per GPU (SURVEY.md 8(d)
recipe, seed 0x5517, resident in HBM before the timed region), one ssim_cuda_compute_device() call = ONE launch of
the persistent fused kernel (per-frame reduction inside), producing FRAMES maps and FRAMES global SSIM values.  Frames are
independent, so with N GPUs each rank owns its own FRAMES frames and there is no data-path collective (weak
scaling); the timed region is bracketed by barrier + synchronize and the slowest rank's device time counts.

Printed JSON (one line, rank 0):
  value      whole-job Mpix/s (all ranks' pixels / max-over-ranks device time), inputs and outputs in HBM
  e2e        same metric through the reference-facing API rmgr_ssim_compute_ssim() with pinned HOST buffers:
             H2D of both images and D2H of the map + scalar inside the timed region, one blocking call per frame,
             the calls issued from --e2e-threads host threads (default 2; the API is re-entrant like the reference's)
  roofline   the fused kernel (the only kernel of a step; CUDA events over the timed region): algorithmic
             230 flop/pixel / duration vs the FP32 FFMA peak 148 SM x 128 lanes x 2 x sm_max_mhz
             (MEASURED_PEAKS.json has no FP32 figure; tools/microbench measured 97-99% of this nominal peak),
             plus the HBM view (6 B/pixel with map vs the measured copy bandwidth)
  cpu_baseline  the UNMODIFIED reference (float build, AUTO dispatch = FMA blur, OpenMP over all host cores),
             compiled from /root/reference into oracle/_ref, timed on a bounded sample of the same frames
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 3840, 2160
FRAMES = 64                      # 4K pairs per GPU per step: 1.06 GB of inputs + 2.1 GB of maps (>> 126 MB L2)
FLOP_PER_PIXEL = 230.0           # algorithmic work of the separable formulation
BYTES_PER_PIXEL_MAP = 6.0        # 2 B read + 4 B map write
METRIC = "ssim_mpix_per_s_4k_with_map"
UNIT = "Mpix/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary BASELINE.json configs (1080p, 16384^2 strips, 1080p batch)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads issuing the blocking reference-API calls of the e2e leg")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), float(p["sm_max_mhz"]), "measured"
    return 6650.0, 1965.0, "fallback"          # B200_PROFILING.md fallback


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def pin_to_gpu_numa_node(index):
    """Binds this process to the CPUs of the NUMA node the GPU hangs off (so that pinned staging memory, allocated afterwards,
    is node-local).  Returns a short description for the bench line; silently does nothing where the topology is flat."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                  # nvml prints an 8-digit domain, sysfs a 4-digit one
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        if node < 0 or len(nodes) < 2:
            return "flat topology (%d NUMA node(s)), no binding" % len(nodes)
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "bound to NUMA node %d (%d CPUs)" % (node, len(cpus))
        return "NUMA node %d has no allowed CPUs, no binding" % node
    except Exception as e:                                 # noqa: BLE001 -- topology files differ between hosts
        return "no binding (%s)" % type(e).__name__


def pcie_probe(torch, dist, world, dev, seconds=0.15):
    """Concurrent H2D + D2H copy rates of THIS rank while every other rank does the same (barrier first): the PCIe / host
    memory ceiling of the e2e leg at this N.  Returns (h2d GB/s, d2h GB/s) of this rank."""
    n = 64 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    both()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    reps = 0
    t0 = time.perf_counter()
    e[0].record(s1); e[2].record(s2)
    while time.perf_counter() - t0 < seconds or reps < 4:
        both()
        reps += 1
    e[1].record(s1); e[3].record(s2)
    torch.cuda.synchronize()
    return n * reps / e[0].elapsed_time(e[1]) / 1e6, n * reps / e[2].elapsed_time(e[3]) / 1e6


def c_client(w, h, with_map, reps=50):
    """The same latency measured by a compiled C++ caller (ssim_b200/bin/latency_client, csrc/latency_client.cpp): no Python in
    the timed path.  Returns its JSON, or a note when the binary is not built."""
    import subprocess
    exe = os.path.join(ROOT, "ssim_b200", "bin", "latency_client")
    if not os.path.exists(exe):
        return {"unavailable": "ssim_b200/bin/latency_client not built"}
    try:
        r = subprocess.run([exe, str(w), str(h), str(with_map), str(reps)], capture_output=True, text=True, timeout=120)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:                                 # noqa: BLE001
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def config(args, n):
    return {"workload": "3840x2160 8-bit grayscale pairs with per-pixel map (BASELINE.json configs[2]), %d pairs per GPU per step" % args.frames,
            "frames_per_gpu_per_step": args.frames, "width": W, "height": H,
            "l2": "per-step working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (args.frames * W * H * 6 / 1e9),
            "parallelism": "frames sharded one batch per GPU (dp%d), no data-path collective" % n}


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def host_threads_for_reference():
    """The reference arm gets ALL host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to every rank, which
    would silently run the reference's OpenMP pool on one thread: the OpenMP runtime the reference library is linked with
    (libgomp) is told explicitly, and what it then reports is what goes into `cores`."""
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    want = max(1, min(want, 64))                  # the reference caps its pool at 64 threads (src/ssim.cpp:1025-1029)
    try:
        gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
        gomp.omp_set_num_threads(C.c_int(want))
        gomp.omp_get_max_threads.restype = C.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return want


def time_reference(frames_host, seconds=None, steps=None, warmup=1):
    """Runs the unmodified reference (oracle/_ref/libref_f32.so, rmgr_ssim_compute_ssim_openmp) on host frames.
    Either for ~`seconds` of wall clock or for `steps` timed steps of len(frames_host) frames each."""
    import numpy as np

    import oracle
    from ssim_b200._abi import make_params
    lib = oracle.ref_lib("f32")
    lib.ref_select_impl(oracle.IMPL_AUTO)
    cores = host_threads_for_reference()
    m = np.empty((H, W), dtype=np.float32)
    out = C.c_float()

    def one(a, b):
        p = make_params(a, b, W, H, ssim_map=m)
        rc = lib.rmgr_ssim_compute_ssim_openmp(C.byref(out), C.byref(p))
        assert rc == 0, rc

    for _ in range(warmup):
        for a, b in frames_host:
            one(a, b)
    done = 0
    t0 = time.perf_counter()
    if steps is not None:
        for _ in range(steps):
            for a, b in frames_host:
                one(a, b)
                done += 1
    else:
        while time.perf_counter() - t0 < seconds or done < 3:
            a, b = frames_host[done % len(frames_host)]
            one(a, b)
            done += 1
    dt = time.perf_counter() - t0
    return done * W * H / dt / 1e6, cores, done, dt, float(out.value)


def time_port(seconds):
    """Fallback when oracle/_ref is absent: the plain-C double restatement (oracle/liboracle.so, OpenMP over rows) on 960x540
    crops of the synthetic 4K pairs.  Much slower than the reference's SIMD path; kind = "port"."""
    import numpy as np

    import oracle
    from ssim_b200.synth import synth_pair
    w, h = 960, 540
    a, b = synth_pair(W, H, 0)
    a = np.ascontiguousarray(a[:h, :w]); b = np.ascontiguousarray(b[:h, :w])
    cores = oracle.oracle_lib().ssim_oracle_num_threads()
    oracle.oracle_ssim(a, b, want_map=True)
    done = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds or done < 2:
        s, _, _ = oracle.oracle_ssim(a, b, want_map=True)
        done += 1
    dt = time.perf_counter() - t0
    return done * w * h / dt / 1e6, cores, done, dt, float(s)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from ssim_b200.synth import synth_pair
    n = args.gpus
    if not oracle.have_ref():
        # the reference could not be compiled into oracle/_ref on the build host: time the restatement instead
        per = []
        for _ in range(args.warmup + args.steps):
            per.append(time_port(2.0))
        per = per[args.warmup:]
        mpix = sum(p[0] for p in per) / len(per)
        line = {"impl": "reference", "metric": METRIC, "value": round(mpix, 2), "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1e3 * sum(p[3] for p in per) / len(per), 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config(args, n),
                "cpu_baseline": {"value": round(mpix, 2), "unit": UNIT, "cores": per[0][1], "cpu_model": cpu_model(), "kind": "port",
                                 "sample": "oracle/liboracle.so (plain-C double restatement, OpenMP) on 960x540 crops of the synthetic 4K pairs with map, 2 s per step"},
                "e2e": {"value": round(mpix, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    sample = 8                                     # frames per step: bounded sample of the 64-frame batch (~40 ms of CPU per step)
    frames = [synth_pair(W, H, f) for f in range(sample)]
    # untimed spin-up (thread pool, page faults, CPU clocks): a cold 0.2 s run measured 1.0 k Mpix/s where the warm
    # steady state of the same call is 1.8-1.9 k on the 16-core box; the reference deserves its steady state
    time_reference(frames[:4], seconds=3.0)
    mpix, cores, done, dt, _ = time_reference(frames, steps=args.steps, warmup=max(1, min(args.warmup, 3)))
    cfg = config(args, n)
    cfg["reference_arm_step"] = "a step of this arm is a bounded sample of the workload: %d of the %d pairs (ms_per_step is per sample step; value is a rate)" % (sample, args.frames)
    cfg["sample_frames_per_step"] = sample
    line = {"impl": "reference", "metric": METRIC, "value": round(mpix, 2), "unit": UNIT, "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": round(mpix, 2), "unit": UNIT, "cores": cores, "cpu_model": cpu_model(), "kind": "reference",
                             "sample": "%d synthetic 4K pairs with map per step (frames 0..%d), %d steps, rmgr_ssim_compute_ssim_openmp of the unmodified float build" % (sample, sample - 1, args.steps)},
            "e2e": {"value": round(mpix, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML every ~5 ms while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.period = 0.005
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ secondary configs
def run_extras(args, api, torch, dist, local, rank, world, stream, barrier):
    """BASELINE.json configs[1], [3], [4], measured the same way (device-timed, max over ranks); informational."""
    from ssim_b200 import parallel
    dev = torch.device("cuda", local)
    sh = stream.cuda_stream
    out = {}

    def timed(fn, iters, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def one_call_us(fn, n=20):
        """events around ONE call with an idle stream before it: the latency a caller sees, host-side launch cost included"""
        ts = []
        for i in range(n + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts), min(ts)

    # configs[1]: one 1920x1080 pair, global SSIM only (no map)
    w, h = 1920, 1080
    a = torch.empty((h, w), dtype=torch.uint8, device=dev)
    b = torch.empty((h, w), dtype=torch.uint8, device=dev)
    sums = torch.empty(1, dtype=torch.float64, device=dev)
    val = torch.empty(1, dtype=torch.float32, device=dev)
    api.synth_fill(local, sh, a.data_ptr(), w, b.data_ptr(), w, w, h, 0, 0)
    f1080 = lambda: api.compute_device(local, sh, w, h, 0, h, 1, a.data_ptr(), w, 0, b.data_ptr(), w, 0, None, 0, 0, sums.data_ptr(), val.data_ptr())  # noqa: E731
    ms = timed(f1080, 50)
    med, mn = one_call_us(f1080)
    out["1080p_pair_no_map"] = {"us_per_pair": round(med, 2), "us_per_pair_min": round(mn, 2), "us_per_pair_queued": round(ms * 1e3, 2),
                                "mpix_per_s": round(w * h / med, 1), "ssim": float(val.item()),
                                "note": "one pair per call = one launch, on every rank: us_per_pair = events around a single call on an idle stream "
                                        "(median of 20), us_per_pair_queued = 50 calls queued back to back (L2-resident inputs)"}
    # the same config end to end: host buffers through rmgr_ssim_compute_ssim, no map (H2D of both images + the scalar back)
    ha, hb = a.cpu().pin_memory(), b.cpu().pin_memory()
    na, nb = ha.numpy(), hb.numpy()
    for _ in range(3):
        api.compute_ssim(na, nb)
    barrier()
    t0 = time.perf_counter()
    reps = 30
    for _ in range(reps):
        e2e_val, _ = api.compute_ssim(na, nb)
    dt = (time.perf_counter() - t0) / reps
    out["1080p_pair_no_map"]["e2e_us_per_pair"] = round(dt * 1e6, 1)
    out["1080p_pair_no_map"]["e2e_mpix_per_s"] = round(w * h / dt / 1e6, 1)
    out["1080p_pair_no_map"]["e2e_ssim"] = float(e2e_val)
    if rank == 0:
        out["1080p_pair_no_map"]["c_client"] = c_client(1920, 1080, 0)

    # configs[3]: ONE 16384x16384 pair split into row strips with 5-row halos across the ranks, map strips written,
    # double partial sums all-reduced with NCCL inside the timed region
    W16 = 16384
    s0, s1, oy, orows = parallel.strip_bounds(W16, world, rank)
    a = torch.empty((s1 - s0, W16), dtype=torch.uint8, device=dev)
    b = torch.empty((s1 - s0, W16), dtype=torch.uint8, device=dev)
    m = torch.empty((orows, W16), dtype=torch.float32, device=dev)
    api.synth_fill(local, sh, a.data_ptr(), W16, b.data_ptr(), W16, W16, s1 - s0, s0, 0)

    def strips():
        api.compute_device(local, sh, W16, s1 - s0, oy, orows, 1, a.data_ptr(), W16, 0, b.data_ptr(), W16, 0, m.data_ptr(), W16, 0, sums.data_ptr(), None)
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)

    ms = timed(strips, 10)
    strips()
    torch.cuda.synchronize()
    out["16384x16384_strips_with_map"] = {"ms": round(ms, 4), "mpix_per_s": round(W16 * W16 / ms / 1e3, 1), "scaling": "strong",
                                          "ssim": float(parallel.mean_from_partials(float(sums.item()), W16, W16)),
                                          "collective": "none (1 GPU)" if world == 1 else "NCCL all-reduce of 1 double per step"}
    if world > 1:
        # same strips, the cross-GPU sum fused into the kernel: the last warp of every rank's launch stores its strip sum into every
        # peer's exchange buffer over NVLink and adds up what lands in its own (ssim_cuda_compute_strip_allreduce); the
        # buffers of the other processes are mapped through CUDA IPC handles
        buf, handle = api.exchange_create(local)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        peers = [buf if r == rank else api.exchange_open(local, handles[r]) for r in range(world)]
        allsum = torch.zeros(1, dtype=torch.float64, device=dev)
        allval = torch.zeros(1, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        epoch = [0]

        def strips_p2p():
            epoch[0] += 1
            api.compute_strip_allreduce(local, sh, W16, s1 - s0, oy, orows, W16, a.data_ptr(), W16, b.data_ptr(), W16, m.data_ptr(), W16,
                                        peers, rank, epoch[0], allsum.data_ptr(), allval.data_ptr(), status.data_ptr())

        ms2 = timed(strips_p2p, 10)
        torch.cuda.synchronize()
        out["16384x16384_strips_with_map_peer_memory"] = {"ms": round(ms2, 4), "mpix_per_s": round(W16 * W16 / ms2 / 1e3, 1), "scaling": "strong",
                                                          "ssim": float(allval.item()), "status": int(status.item()),
                                                          "collective": "strip sums exchanged by NVLink peer stores inside the fused kernel (one launch per rank, no NCCL call)"}
        dist.barrier()
        del a, b, m
        # strong-scaling efficiency needs the 1-GPU time of the same image on the same box: every rank times the whole image alone
        a = torch.empty((W16, W16), dtype=torch.uint8, device=dev)
        b = torch.empty((W16, W16), dtype=torch.uint8, device=dev)
        m = torch.empty((W16, W16), dtype=torch.float32, device=dev)
        api.synth_fill(local, sh, a.data_ptr(), W16, b.data_ptr(), W16, W16, W16, 0, 0)
        ms1 = timed(lambda: api.compute_device(local, sh, W16, W16, 0, W16, 1, a.data_ptr(), W16, 0, b.data_ptr(), W16, 0, m.data_ptr(), W16, 0, sums.data_ptr(), None), 5)
        for key, t in (("16384x16384_strips_with_map", ms), ("16384x16384_strips_with_map_peer_memory", ms2)):
            out[key]["single_gpu_ms_same_box"] = round(ms1, 4)
            out[key]["strong_scaling_efficiency"] = round(ms1 / (world * t), 4)
    del a, b, m

    # SURVEY 8(e), single-process form: ssim_cuda_compute_strips() drives ALL visible GPUs of the box from rank 0 (host image in,
    # strips + halos copied to every GPU, strip sums exchanged inside the kernels over peer memory), checked against the
    # unmodified reference's value for the same synthetic image (tests/golden/golden.json)
    ngpu = torch.cuda.device_count()
    if rank == 0 and ngpu >= 1:
        import numpy as np
        a = torch.empty((W16, W16), dtype=torch.uint8, device=dev)
        b = torch.empty((W16, W16), dtype=torch.uint8, device=dev)
        api.synth_fill(local, sh, a.data_ptr(), W16, b.data_ptr(), W16, W16, W16, 0, 0)
        ha, hb = a.cpu().pin_memory(), b.cpu().pin_memory()
        del a, b
        hm = torch.empty((W16, W16), dtype=torch.float32).pin_memory()
        devices = list(range(min(ngpu, max(world, 1))))
        with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as fh:
            want = json.load(fh)["synthetic"]["16384x16384_f0"]["ref_f64_auto"]
        import ctypes as C2
        lib = api.cuda_lib()
        got = C2.c_float()
        devs = (C2.c_int * len(devices))(*devices)

        def call():
            rc = lib.ssim_cuda_compute_strips(len(devices), devs, W16, W16, ha.data_ptr(), 1, W16, hb.data_ptr(), 1, W16, hm.data_ptr(), 1, W16, C2.byref(got))
            assert rc == 0, (rc, lib.ssim_cuda_last_error_string())

        call()
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        out["compute_strips_single_process"] = {"devices": devices, "ms_wall": round(dt * 1e3, 2), "mpix_per_s": round(W16 * W16 / dt / 1e6, 1),
                                                "ssim": float(got.value), "reference_f64_ssim": want, "abs_diff": abs(float(got.value) - want),
                                                "map_mean": float(hm.numpy().mean(dtype=np.float64)),
                                                "note": "host image -> strips on all listed GPUs, maps back to the host; wall clock incl. all copies"}
        assert abs(float(got.value) - want) <= 2e-6, out["compute_strips_single_process"]
        del ha, hb, hm
    barrier()

    # all channels of an interleaved 1080p RGB pair (maps included) in one call vs one call per channel
    if rank == 0:
        import numpy as np
        rng = np.random.default_rng(5)
        ra = torch.from_numpy(rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).pin_memory()
        rb = torch.from_numpy(np.clip(ra.numpy().astype(np.int16) + rng.integers(-20, 21, ra.shape), 0, 255).astype(np.uint8)).pin_memory()
        rm = torch.empty((1080, 1920, 3), dtype=torch.float32).pin_memory()
        xa, xb, xm = ra.numpy(), rb.numpy(), rm.numpy()
        lib = api.cuda_lib()
        import ctypes as C2
        outv = (C2.c_float * 3)()

        def all_channels():
            rc = lib.ssim_cuda_compute_channels(local, 1920, 1080, 3, xa.ctypes.data, 1920 * 3, xb.ctypes.data, 1920 * 3, xm.ctypes.data, 1920 * 3, outv)
            assert rc == 0, rc

        def per_channel():
            for ch in range(3):
                api.compute_ssim(xa, xb, width=1920, height=1080, step_a=3, step_b=3, stride_a=5760, stride_b=5760, a_off=ch, b_off=ch,
                                 ssim_map=xm, map_step=3, map_stride=5760, map_off=ch)

        res = {}
        for name, fn in (("one_call_all_channels", all_channels), ("one_call_per_channel", per_channel)):
            for _ in range(2):
                fn()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            res[name + "_ms"] = round((time.perf_counter() - t0) / 10 * 1e3, 3)
        res["ssim_rgb"] = [float(v) for v in outv]
        out["1080p_rgb_pair_with_maps_host_api"] = res
    barrier()

    # configs[4]: 4096 x 1080p pairs with maps, 512 per GPU (weak scaling; fewer per GPU when more than 8 ranks are not available)
    F = 512
    w, h = 1920, 1080
    a = torch.empty((F, h, w), dtype=torch.uint8, device=dev)
    b = torch.empty((F, h, w), dtype=torch.uint8, device=dev)
    m = torch.empty((F, h, w), dtype=torch.float32, device=dev)
    fs = torch.empty(F, dtype=torch.float64, device=dev)
    fv = torch.empty(F, dtype=torch.float32, device=dev)
    for f in range(F):
        api.synth_fill(local, sh, a[f].data_ptr(), w, b[f].data_ptr(), w, w, h, 0, rank * F + f)
    ms = timed(lambda: api.compute_device(local, sh, w, h, 0, h, F, a.data_ptr(), w, w * h, b.data_ptr(), w, w * h, m.data_ptr(), w, w * h,
                                          fs.data_ptr(), fv.data_ptr()), 5)
    out["1080p_batch_512_per_gpu_with_maps"] = {"ms": round(ms, 3), "mpix_per_s": round(world * F * w * h / ms / 1e3, 1), "scaling": "weak",
                                                "ssim_first_last": [float(fv[0].item()), float(fv[-1].item())]}
    del a, b, m

    # 16-bit pixels (L = 65535), 32 x 4K pairs with maps per GPU; pixels = 257 x the 8-bit synthetic
    # frames, so the per-frame SSIM must equal the 8-bit one (scale invariance) -- reported next to it
    F = 32
    w, h = 3840, 2160
    a8 = torch.empty((h, w), dtype=torch.uint8, device=dev)
    b8 = torch.empty((h, w), dtype=torch.uint8, device=dev)
    a = torch.empty((F, h, w), dtype=torch.int16, device=dev)
    b = torch.empty((F, h, w), dtype=torch.int16, device=dev)
    m = torch.empty((F, h, w), dtype=torch.float32, device=dev)
    fv16 = torch.empty(F, dtype=torch.float32, device=dev)
    for f in range(F):
        api.synth_fill(local, sh, a8.data_ptr(), w, b8.data_ptr(), w, w, h, 0, rank * F + f)
        a[f] = (a8.to(torch.int32) * 257).to(torch.int16)          # bit pattern of the uint16 value
        b[f] = (b8.to(torch.int32) * 257).to(torch.int16)
    api.compute_device(local, sh, w, h, 0, h, 1, a8.data_ptr(), w, 0, b8.data_ptr(), w, 0, None, 0, 0, None, val.data_ptr())
    ms = timed(lambda: api.compute_device_u16(local, sh, w, h, 0, h, F, a.data_ptr(), 2 * w, 2 * w * h, b.data_ptr(), 2 * w, 2 * w * h,
                                              m.data_ptr(), w, w * h, None, fv16.data_ptr()), 5)
    torch.cuda.synchronize()
    out["4k_u16_batch_32_per_gpu_with_maps"] = {"ms": round(ms, 3), "mpix_per_s": round(world * F * w * h / ms / 1e3, 1), "scaling": "weak",
                                                "ssim_last_u16": float(fv16[-1].item()), "ssim_last_same_frame_u8": float(val.item())}
    return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from ssim_b200 import api

    n = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == n, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (n, world)
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa_node(local)             # before any pinned allocation: staging memory lands on the GPU's node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = api.cuda_lib()
    assert lib.ssim_cuda_device_count() > local, "no CUDA device: ssim_b200 has no CPU fallback"
    dev = torch.device("cuda", local)
    F = args.frames
    npx = W * H

    # ---- inputs resident in HBM: frames rank*F .. rank*F+F-1 of the synthetic sweep
    dA = torch.empty((F, H, W), dtype=torch.uint8, device=dev)
    dB = torch.empty((F, H, W), dtype=torch.uint8, device=dev)
    dMap = torch.empty((F, H, W), dtype=torch.float32, device=dev)
    dSums = torch.empty(F, dtype=torch.float64, device=dev)
    dSsim = torch.empty(F, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    for f in range(F):
        api.synth_fill(local, sh, dA[f].data_ptr(), W, dB[f].data_ptr(), W, W, H, 0, rank * F + f)
    torch.cuda.synchronize()

    def step():
        api.compute_device(local, sh, W, H, 0, H, F, dA.data_ptr(), W, npx, dB.data_ptr(), W, npx, dMap.data_ptr(), W, npx,
                           dSums.data_ptr(), dSsim.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-timed whole-job throughput
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = args.steps * lib.ssim_cuda_last_launch_count()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n * F * npx * args.steps / (ms_max * 1e-3) / 1e6
    ssim_first = float(dSsim[0].item())

    # ---- roofline: a step IS one launch of the fused kernel (nothing else runs in the timed region), so its average launch
    # duration is this rank's timed region / steps
    kernel_ms = ms / args.steps
    sampler.period = 0.05                               # the device-timed region is over: sample lazily from here on

    # ---- one 4K pair per launch (latency view; BASELINE.md: <= 42.7 us is the 60% target)
    single = []
    ptrs = [(dA[f].data_ptr(), dB[f].data_ptr(), dMap[f].data_ptr()) for f in range(F)]      # no tensor indexing inside the timed calls
    pS, pV = dSums.data_ptr(), dSsim.data_ptr()
    for i in range(3 + 20):
        pa, pb, pm = ptrs[i % F]
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s0.record(stream)
        api.compute_device(local, sh, W, H, 0, H, 1, pa, W, npx, pb, W, npx, pm, W, npx, pS, pV)
        s1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            single.append(s0.elapsed_time(s1) * 1e3)

    torch.cuda.synchronize()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record(stream)
    for i in range(40):
        pa, pb, pm = ptrs[i % F]
        api.compute_device(local, sh, W, H, 0, H, 1, pa, W, npx, pb, W, npx, pm, W, npx, pS, pV)
    q1.record(stream)
    torch.cuda.synchronize()
    single_queued_us = q0.elapsed_time(q1) * 1e3 / 40

    # ---- end to end through the reference-facing API with pinned host buffers (H2D + D2H inside the timed region)
    eF = min(F, 8)                                     # host frames kept pinned (rotated), 8 x 50 MB
    hA = torch.empty((eF, H, W), dtype=torch.uint8).pin_memory()
    hB = torch.empty((eF, H, W), dtype=torch.uint8).pin_memory()
    hMap = torch.empty((eF, H, W), dtype=torch.float32).pin_memory()
    hA.copy_(dA[:eF])
    hB.copy_(dB[:eF])
    torch.cuda.synchronize()
    nA, nB, nM = hA.numpy(), hB.numpy(), hMap.numpy()
    os.environ["SSIM_CUDA_DEVICE"] = str(local)

    # The reference API is re-entrant, and its own arm uses every host core: the step's F blocking calls are
    # issued from E2E_THREADS host threads (ctypes drops the GIL), each frame by exactly one call; concurrent calls run
    # on sibling contexts of the device and hide each other's pipeline fill and drain.
    import threading
    e2e_threads = max(1, args.e2e_threads)

    def e2e_step():
        last = [None] * e2e_threads

        def work(t):
            for f in range(t, F, e2e_threads):
                k = f % eF
                last[t], _ = api.compute_ssim(nA[k], nB[k], ssim_map=nM[k])

        if e2e_threads == 1:
            work(0)
        else:
            ths = [threading.Thread(target=work, args=(t,)) for t in range(e2e_threads)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        return last[(F - 1) % e2e_threads]

    h2d_gbs, d2h_gbs = pcie_probe(torch, dist, world, dev)
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(1):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_last = e2e_step()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    td = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
    e2e_value = n * F * npx * e2e_steps / float(td.item()) / 1e6
    # this rank's full-duplex PCIe bound for the metric (2 B/pixel in, 4 B/pixel out, both directions busy at once), summed over ranks
    pb = torch.tensor([1.0 / max(2.0 / (h2d_gbs * 1e3), 4.0 / (d2h_gbs * 1e3)), h2d_gbs, d2h_gbs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pb, op=dist.ReduceOp.SUM)
    pcie_bound_mpix, h2d_all, d2h_all = (float(v) for v in pb.tolist())

    extras = None if args.no_extras else run_extras(args, api, torch, dist, local, rank, world, stream, barrier)

    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_gbs, sm_max_mhz, peak_kind = peaks()
    fp32_peak = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12          # TFLOP/s
    achieved = FLOP_PER_PIXEL * F * npx / (kernel_ms * 1e-3) / 1e12
    hbm_achieved = BYTES_PER_PIXEL_MAP * F * npx / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes of one launch cannot be measured outside a profiler: quoted from the committed ncu capture of this kernel,
    # with its provenance (scaled by the frame count: the capture's traffic is proportional to it)
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "fused_kernel_dram_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get("frames") and tj.get("dram_bytes_per_launch"):
            traffic = round(tj["dram_bytes_per_launch"] * F / tj["frames"])
            traffic_source = "ncu --set full capture %s (%s, %d frames per launch, kernel %s), scaled to %d frames" % (
                tj.get("capture", "?"), tj.get("captured_at_commit", "?"), tj["frames"], tj.get("kernel", "?"), F)

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(args, n),
            "clocks": sampler.summary(),
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": F * 2 * npx, "d2h_bytes_per_step": F * (npx * 4 + 4),
                    "api": "rmgr_ssim_compute_ssim (librmgr-ssim.so), one blocking call per 4K pair, pinned host buffers, calls issued from %d host thread(s)" % e2e_threads,
                    "steps": e2e_steps,
                    "pcie_bound_mpix": round(pcie_bound_mpix, 1), "frac_of_pcie_bound": round(e2e_value / pcie_bound_mpix, 4),
                    "pcie_probe": {"h2d_gbs_all_ranks": round(h2d_all, 1), "d2h_gbs_all_ranks": round(d2h_all, 1),
                                   "how": "every rank copies 64 MiB pinned buffers H2D and D2H at the same time on two streams, all ranks concurrently; "
                                          "bound = sum over ranks of 1 / max(2 B/px / h2d, 4 B/px / d2h)"},
                    "host_binding": numa},
            "gpu_launches": launches,
            "roofline": {"bound": "fp32", "achieved": round(achieved, 2), "peak": round(fp32_peak, 2), "unit": "TFLOP/s",
                         "frac": round(achieved / fp32_peak, 4), "traffic": traffic, "traffic_source": traffic_source,
                         "algorithmic_bytes": int(BYTES_PER_PIXEL_MAP * F * npx),
                         "kernel": "void ssimk::ssim_fused_kernel<1, false>(CUtensorMap_st, CUtensorMap_st, ssimk::FusedParams, ssimk::ExchangeParams)",
                         "kernel_ms_per_launch": round(kernel_ms, 4),
                         "peak_source": "nominal 148 SM x 128 lanes x 2 x sm_max_mhz (%s MEASURED_PEAKS.json sm_max_mhz); FFMA microbench reaches 97-99%% of it" % peak_kind,
                         "hbm": {"achieved": round(hbm_achieved, 1), "peak": hbm_gbs, "unit": "GB/s", "frac": round(hbm_achieved / hbm_gbs, 4),
                                 "peak_source": "%s copy bandwidth" % peak_kind}},
            "single_pair_us": {"median": round(statistics.median(single), 2), "min": round(min(single), 2),
                               "mpix_per_s_at_median": round(npx / statistics.median(single), 1),
                               "queued_us_per_pair": round(single_queued_us, 2),
                               "note": "one 4K pair with map per call = one launch: median/min = events around a single call on an idle stream "
                                       "(host-side launch cost of the Python/ctypes caller included), queued = 40 calls (rotating frames) queued back to back",
                               "c_client": c_client(3840, 2160, 1)},
            "ssim_frame0": ssim_first, "ssim_e2e_last": float(e2e_last)}
    if extras is not None:
        line["other_configs"] = extras

    if n == 1 and not args.no_cpu_baseline:
        import oracle
        if oracle.have_ref():
            frames = [(nA[k], nB[k]) for k in range(min(eF, 4))]
            mpix, cores, done, cdt, cpu_ssim = time_reference(frames, seconds=args.cpu_seconds)
            line["cpu_baseline"] = {"value": round(mpix, 1), "unit": UNIT, "cores": cores, "cpu_model": cpu_model(), "kind": "reference",
                                    "sample": "%d calls of rmgr_ssim_compute_ssim_openmp (unmodified float build, AUTO=FMA dispatch) on synthetic 4K pairs with map, %.1f s" % (done, cdt),
                                    "ssim_last": cpu_ssim}
        else:
            mpix, cores, done, cdt, cpu_ssim = time_port(min(args.cpu_seconds, 10.0))
            line["cpu_baseline"] = {"value": round(mpix, 2), "unit": UNIT, "cores": cores, "cpu_model": cpu_model(), "kind": "port",
                                    "sample": "%d calls of oracle/liboracle.so (plain-C double restatement, OpenMP) on 960x540 crops of the synthetic 4K pairs with map, %.1f s" % (done, cdt),
                                    "ssim_last": cpu_ssim}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
