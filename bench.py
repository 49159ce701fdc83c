#!/usr/bin/env python
"""Benchmark of the RawFormer inference hot path on B200 (BASELINE.json metric: MP/s of RAW input).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S|B|L] [--variant flca|ml]
                    [--row-tiled]

A "step" = one forward over one synthetic SID-Sony-shaped frame (raw 2848x4256 -> packed 1424x2128x4) per GPU.
N = 1 workload = BASELINE configs[1]: RawFormer-S, full frame, bf16.  N > 1 (torchrun) is image-parallel: every rank
runs its own frames, no data-path collective ("weak" scaling); timing = max over ranks of CUDA-event time.

Prints ONE JSON line (see the driver contract): value = device-resident throughput, e2e = same metric through the
public nn.Module call with pinned HOST buffers (H2D + forward + D2H inside the timed region), roofline = the
dominant kernel of the step (per-launch CUDA events on the launching stream), cpu_baseline = the functional-PyTorch
CPU port of the reference timed on this box's host cores on a bounded sample.
`--impl reference` times that CPU port alone (rank 0 only) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
# one hardware queue per stream: the copy streams of FramePipeline must not alias the compute stream's connection
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

H_RAW, W_RAW = 2848, 4256  # SID Sony frame
MP_FRAME = H_RAW * W_RAW / 1e6
METRIC = "megapixels/sec of RAW input (SID Sony full frame)"
SIZES = {"S": 32, "B": 48, "L": 64}
TENSOR_KERNELS = ("gemm_", "conv3x3_out", "down_conv3x3", "up_convT", "skip_reduce")


def workload_name(args, world):
    return (f"RawFormer-{args.size} ({args.variant}) forward, full SID Sony frame raw {H_RAW}x{W_RAW} "
            f"(packed 1424x2128x4), {args.frames} frame(s) per GPU per step, random-init weights")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_burst": d["bf16_tflops"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_burst": 1590.0, "src": "fallback"}


def bind_to_gpu_cpus(gpu_index):
    """Pin this process to the CPUs NVML reports as local to the GPU (so that pinned host buffers land on the GPU's NUMA
    node and the copy threads run next to it).  Returns a short description for the JSON line; never fails the run."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and w * 64 + b < n]
        avail = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in avail]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} GPU-local CPUs"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"
    return "not bound"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).

    In-process NVML from a background thread, initialised BEFORE the timed region: spawning `nvidia-smi -lms` right at
    the start of the region (driver/NVML start-up next to the launch loop) made one timed run in three 40-60 % slower
    than the kernels' own CUDA-event times.  Falls back to nvidia-smi when pynvml is missing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        import threading

        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.handle = None
        self.mx = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:  # noqa: BLE001
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, rs))
            except Exception:  # noqa: BLE001
                pass
            # sparse on purpose: about one NVML query in fifteen stalled for ~100 ms and stalled the GPU work with it (a
            # 150 ms timed region read 10.1 instead of 7.5 ms/step), so the sampler covers the load window = warm-up steps
            # + timed steps (the same kernels back to back) with a few queries instead of many inside the timed region
            self.stop_flag.wait(0.25)

    def start(self):
        import threading

        if self.nvml is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0}
        if self.thread is None:
            return self._smi_once(out)
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if not self.samples:
            return self._smi_once(out)
        out["sm_mhz"] = statistics.median(s for s, _ in self.samples)
        out["samples"] = len(self.samples)
        seen = set()
        for _, rs in self.samples:
            for name, bit in self.REASONS:
                if rs & bit:
                    seen.add(name)
        out["reasons"] = sorted(seen)
        return out

    def _smi_once(self, out):
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                str(self.gpu)], capture_output=True, text=True, timeout=10)
            f = [t.strip() for t in r.stdout.strip().split(",")]
            out["sm_mhz"], out["sm_max_mhz"], out["samples"] = float(f[0]), float(f[1]), 1
            out["note"] = "single nvidia-smi sample after the timed region (pynvml unavailable)"
        except Exception:  # noqa: BLE001
            pass
        return out


def cpu_port_setup(size, variant, h, w, seed=0):
    import torch

    import rf_testlib as T
    from oracle import rawformer_torch as P

    torch.set_num_threads(os.cpu_count() or 1)
    dim = SIZES[size]
    sd = T.make_state_dict(T.build_model(variant, dim), seed=1234, scale=1.0)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(1, 1, h, w, generator=g)
    return (lambda: P.rawformer_forward(sd, x, variant)), torch.get_num_threads()


def time_cpu_port(size, variant, budget_s=25.0):
    """Bounded CPU sample: 1024x1024 raw crop first; the full frame once if the budget allows."""
    fn, threads = cpu_port_setup(size, variant, 1024, 1024)
    fn()  # warm-up
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    mp = 1024 * 1024 / 1e6
    best = {"value": mp / dt, "sample": f"RawFormer-{size} ({variant}) fp32, 1 frame raw 1024x1024 (1.05 MP), 1 warm-up + 1 timed"}
    est_full = dt * (MP_FRAME / mp) * 2.2  # the reference is ~2x slower per MP at full frame
    if est_full < budget_s:
        fn_full, _ = cpu_port_setup(size, variant, H_RAW, W_RAW)
        t0 = time.perf_counter()
        fn_full()
        dt = time.perf_counter() - t0
        best = {"value": MP_FRAME / dt, "sample": f"RawFormer-{size} ({variant}) fp32, 1 full frame raw {H_RAW}x{W_RAW} (12.12 MP), 1 timed run"}
    best.update(unit="MP/s", cores=threads, kind="port")
    return best


def time_torch_eager_b200(size, variant, dev):
    """The reference's own operator sequence (oracle/rawformer_torch.py = the ATen calls of the reference modules) run
    EAGERLY on the B200: the "existing Blackwell path" of SURVEY 8d.  A reported baseline like cpu_baseline, never a
    product path.  fp32 with TF32 contractions, and the same under torch.autocast(bf16)."""
    import torch

    import rf_testlib as T
    from oracle import rawformer_torch as P

    dim = SIZES[size]
    sd = {k: v.to(dev) for k, v in T.make_state_dict(T.build_model(variant, dim), seed=1234, scale=1.0).items()}
    x = torch.rand(1, 1, H_RAW, W_RAW, generator=torch.Generator().manual_seed(0)).to(dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    out = {}
    for name, ctx in (("fp32_tf32", None), ("autocast_bf16", torch.autocast("cuda", dtype=torch.bfloat16))):
        def step():
            if ctx is None:
                return P.rawformer_forward(sd, x, variant)
            with ctx:
                return P.rawformer_forward(sd, x, variant)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name] = {"value": MP_FRAME / (ms * 1e-3), "unit": "MP/s", "ms_per_frame": ms}
        torch.cuda.empty_cache()
    out["what"] = (f"functional-PyTorch port of the reference forward (same ATen operators, eager, cuDNN/cuBLAS) on this B200, "
                   f"RawFormer-{size} ({variant}), full frame, 2 warm-ups + 3 timed")
    return out


def run_reference(args):
    """Reference arm: the CPU port of the reference forward on this box's host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    h = w = 1024
    fn, threads = cpu_port_setup(args.size, args.variant, h, w)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    mp = h * w / 1e6
    val = args.steps * mp / dt
    sample = (f"CPU port of the reference forward (oracle/rawformer_torch.py, fp32, {threads} threads); each step = "
              f"one raw {h}x{w} crop (1.05 MP) of the frame")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "cpu_sample": f"raw {h}x{w} crop per step"},
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import rf_testlib as T

    import bayer_low_light_image_enhancement_b200 as rf
    from bayer_low_light_image_enhancement_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_cpus(local)   # pinned host buffers are first-touched on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    dim = SIZES[args.size]
    cls = rf.RawFormer if args.variant == "flca" else rf.multilevel.RawFormer
    model = cls(dim=dim, precision=args.precision)
    model.load_state_dict(T.make_state_dict(model, seed=1234, scale=1.0), strict=True)  # random-init weights
    model = model.to(dev).eval()
    if not args.no_graph:
        model.enable_cuda_graphs()      # one graph launch per frame (public API; the per-kernel profile below stays eager)
    B = args.frames
    gen = torch.Generator().manual_seed(rank)
    x_host = torch.rand(B, 1, H_RAW, W_RAW, generator=gen).pin_memory()
    out_host = torch.empty(B, 3, H_RAW, W_RAW).pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    with torch.no_grad():
        # ---- device-resident throughput ----
        # pre-warm: a fresh box needs ~1 s of work before clocks / HBM / lazily loaded kernels settle (the first timed
        # run on a cold GPU was 40 % slower than the next ones); then the W warm-up steps of the contract
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 1.5:
            model(x_dev)
            torch.cuda.synchronize()
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()                       # load window: from the first warm-up step to the end of the timed region
        t_load = time.perf_counter()
        nw = 0
        while nw < args.warmup or time.perf_counter() - t_load < 0.6:
            model(x_dev)
            nw += 1
            if nw % 8 == 0:
                torch.cuda.synchronize()
        barrier()
        lib.rf_reset_launch_count()
        model.graph_kernels_replayed = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host0 = time.perf_counter()
        for _ in range(args.steps):
            out = model(x_dev)
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
        e1.record()
        barrier()
        # kernels of this library executed in the timed region: launched one by one (eager) or as nodes of the replayed graphs
        launches = lib.rf_launch_count() + model.graph_kernels_replayed
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end through the public call with host buffers ----
        # rf.FramePipeline = the package's streaming API: every step copies its frame from pinned host memory to the
        # device, runs the forward and copies the fp32 result back to pinned host memory; the copies of neighbouring
        # steps overlap with the forward (three streams), all K steps' copies are inside the timed region.
        pipe = rf.FramePipeline(model, depth=2)
        outs = [out_host, torch.empty_like(out_host).pin_memory()]
        t_pre = time.perf_counter()
        i = 0
        while i < 3 or time.perf_counter() - t_pre < 1.0:
            pipe.submit(x_host, outs[i & 1])
            i += 1
            if i % 4 == 0:
                pipe.flush()
        pipe.flush()
        barrier()
        f0 = torch.cuda.Event(enable_timing=True)
        f0.record()
        pipe.start_after(f0)
        for i in range(args.steps):
            pipe.submit(x_host, outs[i & 1])
        f1 = pipe.finish_event()
        pipe.flush()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        e2e_check = float(outs[(args.steps - 1) & 1].abs().max())   # the result really is on the host
        assert e2e_check == e2e_check and e2e_check > 0.0

        # ---- per-kernel times (CUDA events around every launch, on the launching stream) ----
        agg = {}
        n_prof = 3
        for _ in range(n_prof):
            _, launches_prof = model.forward_profiled(x_dev)
            for l in launches_prof:
                a = agg.setdefault(l["name"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0})
                a["ms"] += l["ms"]
                a["bytes"] += l["bytes"]
                a["flops"] += l["flops"]
                a["n"] += 1
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    total_prof_ms = sum(a["ms"] for a in agg.values())
    top_name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    is_tensor = top_name.startswith(TENSOR_KERNELS)
    if is_tensor:
        achieved = top["flops"] / (top["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops"]}
    else:
        achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"]}
    # DRAM traffic of that kernel per launch, from the committed ncu capture of the same workload (null if not captured)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", f"traffic_{args.size}_{args.precision}.json")
    if args.variant == "flca" and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(top_name, {}).get("dram_bytes_per_launch")
    roof.update(kernel=top_name, launches_per_step=top["n"] // n_prof, share_of_step=top["ms"] / total_prof_ms,
                avg_launch_ms=top["ms"] / top["n"], peak_source=pk["src"], traffic=traffic,
                algorithmic_per_launch=(top["flops"] if is_tensor else top["bytes"]) / top["n"])
    breakdown = sorted(((k, v["ms"] / n_prof) for k, v in agg.items()), key=lambda kv: -kv[1])

    frames_total = args.steps * B * world
    value = frames_total * MP_FRAME / (ms_total * 1e-3)
    e2e_val = frames_total * MP_FRAME / (ms_e2e * 1e-3)
    cpu = time_cpu_port(args.size, args.variant) if world == 1 and not args.no_cpu else None
    eager = None
    if world == 1 and args.torch_eager:
        del model, pipe, x_dev
        torch.cuda.empty_cache()
        eager = time_torch_eager_b200(args.size, args.variant, dev)
    line = {
        "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "precision": args.precision,
                   "parallelism": f"image-parallel x{world}" if world > 1 else "single GPU",
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2, no explicit flush",
                   "launch": "eager" if args.no_graph else "one CUDA graph per frame (model.enable_cuda_graphs)"},
        "e2e": {"value": e2e_val, "unit": "MP/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": ms_e2e / args.steps,
                "api": "FramePipeline.submit (H2D, forward, D2H of neighbouring steps overlapped on three streams)",
                "host_binding": numa},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "clocks": clocks,
        "roofline": roof,
        "kernel_ms_per_step": {k: round(v, 4) for k, v in breakdown},
        "model_roofline": {"flops_per_frame": 320.5 * dim * dim * 3030272 + 875.0 * dim * 3030272,
                           "achieved_tflops": (320.5 * dim * dim + 875.0 * dim) * 3030272 * frames_total / (ms_total * 1e-3) / 1e12},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if eager is not None:
        line["eager_b200_baseline"] = eager
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_rowtiled(args):
    """BASELINE config 4: ONE frame per step, cut into row bands over the N ranks (strong scaling).  Every rank holds the
    whole raw frame; halo rows and the per-image reductions cross the GPUs inside kernels of this library (peer-mapped
    memory over NVLink).  N = 1 runs the same band code with a single band."""
    import torch
    import torch.distributed as dist

    import rf_testlib as T

    import bayer_low_light_image_enhancement_b200 as rf
    from bayer_low_light_image_enhancement_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_cpus(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    dim = SIZES[args.size]
    model = rf.RawFormer(dim=dim, precision="bf16")
    model.load_state_dict(T.make_state_dict(model, seed=1234, scale=1.0), strict=True)
    model = model.to(dev).eval()
    x_host = torch.rand(1, 1, H_RAW, W_RAW, generator=torch.Generator().manual_seed(0)).pin_memory()   # same frame on all ranks
    x_dev = x_host.to(dev)
    if world > 1:
        tiled = rf.RowTiledRawFormer.from_process_group(model, H_RAW, W_RAW)
    else:                                         # one band = the whole frame, same band code, no peers
        from bayer_low_light_image_enhancement_b200.rowtiled import _CommRegion

        region = _CommRegion(rf.RowTiledRawFormer.comm_bytes(model, H_RAW, W_RAW, 1), dev)
        tiled = rf.RowTiledRawFormer(model, H_RAW, W_RAW, 0, 1, [region.ptr], own_region=region)
    band_host = torch.empty(1, 3, tiled.rows, W_RAW).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        # parity of the decomposition (untimed): bands gathered on rank 0 against the whole-frame forward of the same engine
        full = tiled.gather(tiled(x_dev), dst=0) if world > 1 else tiled(x_dev).clone()
        tiled.status()
        parity = None
        if rank == 0:
            whole = model(x_dev)
            rng = float(whole.max() - whole.min())
            mse = float(((full - whole).double() ** 2).mean())
            parity = {"psnr_vs_whole_frame_db": 99.0 if mse == 0 else 10.0 * __import__("math").log10(rng * rng / mse),
                      "max_abs_over_range": float((full - whole).abs().max()) / rng}
            del whole
        del full
        barrier()
        t_pre = time.perf_counter()
        n_pre = torch.zeros(1, device=dev)
        while True:                               # every rank must run the same number of forwards: agree on when to stop
            tiled(x_dev)
            n_pre.fill_(1.0 if time.perf_counter() - t_pre < 1.5 else 0.0)
            if world > 1:
                dist.all_reduce(n_pre, op=dist.ReduceOp.MIN)
            if float(n_pre.item()) == 0.0:
                break
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        lib.rf_reset_launch_count()
        if not args.no_graph:
            tiled.enable_cuda_graphs()            # this rank's band forward as one graph launch per frame
        for _ in range(max(args.warmup, 8)):
            tiled(x_dev)
        barrier()
        per_frame = lib.rf_launch_count()         # graph mode: kernels captured once = kernel nodes replayed per frame
        lib.rf_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            band = tiled(x_dev)
        e1.record()
        barrier()
        launches = lib.rf_launch_count() if args.no_graph else per_frame * args.steps
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None
        tiled.status()
        # end to end: whole frame host -> every rank, forward, this rank's band -> host, every step, one stream
        for _ in range(2):
            x_dev.copy_(x_host, non_blocking=True)
            band_host.copy_(tiled(x_dev), non_blocking=True)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            x_dev.copy_(x_host, non_blocking=True)
            band_host.copy_(tiled(x_dev), non_blocking=True)
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        assert float(band_host.abs().max()) > 0.0
        # per-kernel times of this rank's band (sync-point kernels include the wait for the peers)
        agg = {}
        n_prof = 3
        for _ in range(n_prof):
            _, lp = tiled.forward_profiled(x_dev)
            for l in lp:
                a = agg.setdefault(l["name"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0})
                a["ms"] += l["ms"]; a["bytes"] += l["bytes"]; a["flops"] += l["flops"]; a["n"] += 1
        tiled.status()
        barrier()
    tiled.close()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    pk = peaks()
    compute = {k: v for k, v in agg.items() if not k.startswith("band_")}
    top_name, top = max(compute.items(), key=lambda kv: kv[1]["ms"])
    total_prof_ms = sum(a["ms"] for a in agg.values())
    if top_name.startswith(TENSOR_KERNELS):
        achieved = top["flops"] / (top["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"]}
    else:
        achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"]}
    roof.update(kernel=top_name, launches_per_step=top["n"] // n_prof, share_of_step=top["ms"] / total_prof_ms,
                avg_launch_ms=top["ms"] / top["n"], peak_source=pk["src"], traffic=None, scope="rank 0's band")
    line = {
        "metric": METRIC, "value": args.steps * MP_FRAME / (ms_total * 1e-3), "unit": "MP/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 8), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"RawFormer-{args.size} (flca) forward, ONE SID Sony frame raw {H_RAW}x{W_RAW} per step, "
                               f"row-tiled over {world} GPU(s) (bands of {[r // 16 for _, r in rf.plan_bands(H_RAW, world)]} x 16 rows), "
                               "random-init weights",
                   "precision": "bf16", "parallelism": f"row-tiled x{world}: 4-row halo exchange + all-reduce of the per-image "
                   "reductions per Conv_Transformer, peer-mapped memory over NVLink, no NCCL on the data path",
                   "l2": "per-step working set >> 126 MB L2, no explicit flush",
                   "launch": "eager" if args.no_graph else "one CUDA graph per band per frame"},
        "e2e": {"value": args.steps * MP_FRAME / (ms_e2e * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": int(x_host.numel() * 4) * world,
                "d2h_bytes_per_step": 3 * H_RAW * W_RAW * 4, "ms_per_step": ms_e2e / args.steps,
                "api": "RowTiledRawFormer.forward with pinned host buffers: whole frame H2D on every rank, forward, band D2H",
                "host_binding": numa},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "parity": parity,
        "kernel_ms_per_step": {k: round(v["ms"] / n_prof, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
        "sync_ms_per_step": round(sum(v["ms"] for k, v in agg.items() if k.startswith("band_")) / n_prof, 4),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="S", choices=list(SIZES))
    ap.add_argument("--variant", default="flca", choices=["flca", "ml"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--frames", type=int, default=1, help="frames per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of one CUDA graph per frame")
    ap.add_argument("--torch-eager", action="store_true",
                    help="also time the reference's operator sequence run eagerly by PyTorch on this B200 (baseline leg)")
    ap.add_argument("--row-tiled", action="store_true",
                    help="BASELINE config 4: one frame per step cut into row bands over the ranks (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.row_tiled:
        run_rowtiled(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
