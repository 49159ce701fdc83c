#!/usr/bin/env python
"""Benchmark of the RawFormer inference hot path on B200 (BASELINE.json metric: MP/s of RAW input).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S|B|L] [--variant flca|ml]
                    [--row-tiled]

A "step" = one forward over one synthetic SID-Sony-shaped frame (raw 2848x4256 -> packed 1424x2128x4) per GPU.
N = 1 workload = BASELINE configs[1]: RawFormer-S, full frame, bf16.  N > 1 (torchrun) is image-parallel: every rank
runs its own frames, no data-path collective ("weak" scaling); timing = max over ranks of CUDA-event time.

Prints ONE JSON line (see the driver contract): value = device-resident throughput; e2e = the same metric through the
package's streaming API with pinned HOST buffers in the caller's wire formats -- uint16 sensor frame in (normalised on
the device, WFB/load_dataset.py:88-89), uint8 HWC image out (test.py:117-120) -- H2D + forward + D2H inside the timed
region (e2e_fp32_wires: the same with fp32 tensors both ways); roofline = the dominant CUDA symbol of the step
(per-launch CUDA events on the launching stream); parity_db = PSNR of the timed frame's output against the reference's
CPU forward; cpu_baseline = the functional-PyTorch CPU port of the reference on this box's host cores, one full frame.
N = 1 also reports ms/frame of RawFormer-B / -L and the multi-level variant ("sizes"); N > 1 adds BASELINE config 3
(RawFormer-B image-parallel) and config 4 (RawFormer-L, ONE frame row-tiled over the N GPUs) as sub-records.
`--impl reference` times the CPU port alone on the full frame (rank 0 only) and prints the same line with "impl": "reference".
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
TENSOR_KERNELS = ("gemm_", "conv3x3_out", "conv3x3_lc", "down_conv3x3", "up_convT", "skip_reduce")
# logical launch name -> CUDA symbol family (template instantiation) it runs as, and which roofline bounds that family at
# the benchmarked sizes: the LayerNorm-folded GEMM is ONE symbol although it appears under two logical names
SYMBOL_OF = {
    "gemm_qkv": ("k_tc_gemm<ROWS, LN-fold>", "hbm"), "gemm_pw1": ("k_tc_gemm<ROWS, LN-fold>", "hbm"),
    "gemm_proj_resid": ("k_tc_gemm<ROWS, residual, stats>", "hbm"), "gemm_pw2_resid": ("k_tc_gemm<ROWS, residual>", "hbm"),
    "gemm_cat_reduce": ("k_tc_gemm<ROWS>", "hbm"), "skip_reduce": ("k_tc_gemm<ROWS, stats>", "hbm"),
    "up_convT": ("k_tc_gemm<CONVT>", "hbm"), "conv3x3_out": ("k_tc_gemm<ROWS, lrelu, conv3x3>", "tensor"),
    "down_conv3x3": ("k_tc_gemm<UNSHUFFLE, conv3x3>", "tensor"), "head": ("k_tc_gemm<HEAD, conv3x3>", "hbm"),
    "gemm_pyr_res1": ("k_tc_gemm<ROWS, relu>", "hbm"), "gemm_pyr_res2": ("k_tc_gemm<ROWS, residual, tanh>", "hbm"),
    "dw_qkv_gram": ("k_dw_tma<1>", "hbm"), "dw_gelu": ("k_dw_tma<2>", "hbm"),
    # norm -> 1x1 -> depthwise 3x3 as ONE dense 3x3 conv on the tensor cores (rf_lnconv.cu; C = 32 / 64): bound by the tensor
    # pipe BY DESIGN -- it executes 9x the FLOPs of the 1x1 + depthwise it replaces so that the hidden / q|k tensors never
    # reach HBM; `achieved` counts the EXECUTED dense-conv FLOPs, `algorithmic_hbm` in the record is the compulsory-byte view
    "ffn_fused": ("k_lnconv<FFN>", "tensor"), "qkv_fused": ("k_lnconv<QKV | QK | V>", "tensor"),
    "conv3x3_lc": ("k_lnconv<CONV>", "tensor"), "flca_mod": ("k_im2col_tc<0>", "hbm"), "embed": ("k_im2col_tc<1>", "hbm"),
    "pyr_spatial": ("k_im2col_tc<2|3>", "hbm"), "gemm_gram": ("k_tc_gram", "hbm"),
}


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


def synthetic_u16_frames(batch, seed):
    """Synthetic SID-Sony-shaped sensor frames: uint16 [B,H,W] with black level 512, such that the reference
    normalisation at exposure ratio 100 (WFB/load_dataset.py:88-89) maps them to [0, 1)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    return (512 + torch.randint(0, 159, (batch, H_RAW, W_RAW), generator=g, dtype=torch.int32)).to(torch.uint16)


PRE = {"black": 512.0, "white": 16383.0, "ratio": 100.0, "clamp": True}


def normalise_cpu(u16):
    """The reference loader's arithmetic on the host (fp32, numpy semantics) -> float32 [B,1,H,W]."""
    import numpy as np
    import torch

    from oracle import rawformer_oracle as O

    x = O.preprocess_u16(u16.numpy(), PRE["black"], PRE["white"], PRE["ratio"], PRE["clamp"])
    return torch.from_numpy(np.ascontiguousarray(x[:, None]))


def cpu_port_setup(size, variant, x):
    import torch

    import rf_testlib as T
    from oracle import rawformer_torch as P

    torch.set_num_threads(os.cpu_count() or 1)
    dim = SIZES[size]
    sd = T.make_state_dict(T.build_model(variant, dim), seed=1234, scale=1.0)
    return (lambda: P.rawformer_forward(sd, x, variant)), torch.get_num_threads()


def time_cpu_port(size, variant, x_full):
    """cpu_baseline: the CPU port of the reference forward on ONE full frame (the frame the GPU legs time): one timed
    run after a warm-up on a raw 512x512 crop (thread pool, MKL-DNN primitives).  Returns (record, output)."""
    fn_small, threads = cpu_port_setup(size, variant, x_full[:, :, :512, :512].contiguous())
    fn_small()
    fn, _ = cpu_port_setup(size, variant, x_full)
    t0 = time.perf_counter()
    out = fn()
    dt = time.perf_counter() - t0
    rec = {"value": MP_FRAME / dt, "unit": "MP/s", "cores": threads, "kind": "port",
           "sample": f"RawFormer-{size} ({variant}) fp32 CPU port of the reference forward (oracle/rawformer_torch.py: the "
                     f"reference's ATen operators; einops rearrange -> permute + F.layer_norm), 1 full frame raw "
                     f"{H_RAW}x{W_RAW} (12.12 MP), 1 timed run, {threads} threads"}
    return rec, out


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
    """Reference arm: the CPU port of the reference forward on this box's host cores, rank 0 only, the FULL frame per
    step (same config as the GPU arm: RawFormer-S takes ~6 s per frame on 16 cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x = normalise_cpu(synthetic_u16_frames(1, 0))
    fn, threads = cpu_port_setup(args.size, args.variant, x)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    val = args.steps * MP_FRAME / dt
    sample = (f"CPU port of the reference forward (oracle/rawformer_torch.py, fp32, {threads} threads: the reference's ATen "
              f"operators, einops rearrange -> permute + F.layer_norm); each step = one FULL frame raw {H_RAW}x{W_RAW} (12.12 MP)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "cpu_sample": f"full frame raw {H_RAW}x{W_RAW} per step"},
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Env:
    """Process / device context shared by the legs of one invocation."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_to_gpu_cpus(self.local)   # pinned host buffers are first-touched on the GPU's own NUMA node
        self.dist = dist
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch

        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        import torch

        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def build_model(env, size, variant, precision, graphs=True):
    import rf_testlib as T

    import bayer_low_light_image_enhancement_b200 as rf

    cls = rf.RawFormer if variant == "flca" else rf.multilevel.RawFormer
    model = cls(dim=SIZES[size], precision=precision)
    model.load_state_dict(T.make_state_dict(model, seed=1234, scale=1.0), strict=True)  # random-init weights
    model = model.to(env.dev).eval()
    if graphs:
        model.enable_cuda_graphs()      # one graph launch per frame (public API; the per-kernel profile stays eager)
    return model


def time_resident(env, model, x_dev, steps, warmup, prewarm_s=1.5, load_s=0.6, sampler=None):
    """Device-resident forward: W warm-ups, then exactly `steps` timed steps between barriers; CUDA events, max over ranks."""
    import torch

    from bayer_low_light_image_enhancement_b200 import _lib

    lib = _lib.load()
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < prewarm_s:       # a fresh box needs ~1 s of work before clocks / HBM settle
        model(x_dev)
        torch.cuda.synchronize()
    env.barrier()
    if sampler is not None and env.rank == 0:
        sampler.start()                                   # load window: first warm-up step .. end of the timed region
    t_load = time.perf_counter()
    nw = 0
    while nw < warmup or time.perf_counter() - t_load < load_s:
        model(x_dev)
        nw += 1
        if nw % 8 == 0:
            torch.cuda.synchronize()
    env.barrier()
    lib.rf_reset_launch_count()
    model.graph_kernels_replayed = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(steps):
        out = model(x_dev)
    host_ms = (time.perf_counter() - t_host0) * 1e3 / steps
    e1.record()
    env.barrier()
    launches = lib.rf_launch_count() + getattr(model, "graph_kernels_replayed", 0)
    return env.max_over_ranks(e0.elapsed_time(e1)), int(launches), host_ms, out


def time_pipeline(env, model, x_host, out_hosts, steps, **pipe_kw):
    """End to end through rf.FramePipeline: every step copies its frame from pinned host memory to the device, runs the
    forward and copies the result back to pinned host memory; copies of neighbouring steps overlap with the forward
    (three streams); all K steps' copies are inside the timed region."""
    import torch

    import bayer_low_light_image_enhancement_b200 as rf

    pipe = rf.FramePipeline(model, depth=2, **pipe_kw)
    t_pre = time.perf_counter()
    i = 0
    while i < 3 or time.perf_counter() - t_pre < 1.0:
        pipe.submit(x_host, out_hosts[i & 1])
        i += 1
        if i % 4 == 0:
            pipe.flush()
    pipe.flush()
    env.barrier()
    f0 = torch.cuda.Event(enable_timing=True)
    f0.record()
    pipe.start_after(f0)
    for i in range(steps):
        pipe.submit(x_host, out_hosts[i & 1])
    f1 = pipe.finish_event()
    pipe.flush()
    env.barrier()
    ms = env.max_over_ranks(f0.elapsed_time(f1))
    return ms, pipe


def roofline_records(agg, n_prof, pk, size, precision, variant):
    """roofline = the dominant CUDA SYMBOL of the step (launches grouped by the template instantiation they run as);
    roofline_logical = the dominant logical launch name (round-1 definition, kept for continuity)."""
    total_ms = sum(a["ms"] for a in agg.values())

    def record(ms, by, fl, n, bound, name):
        if bound == "tensor":
            ach = fl / (ms * 1e-3) / 1e12
            r = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"]}
        else:
            ach = by / (ms * 1e-3) / 1e9
            r = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"]}
        r.update(kernel=name, launches_per_step=n // n_prof, share_of_step=ms / total_ms, avg_launch_ms=ms / n,
                 peak_source=pk["src"], algorithmic_per_launch=(fl if bound == "tensor" else by) / n)
        if name.startswith("k_lnconv") and "CONV" not in name:
            gbs = by / (ms * 1e-3) / 1e9
            r["flops_counted"] = "executed (dense 3x3 form = 9x the FLOPs of the 1x1 conv + depthwise 3x3 it replaces)"
            r["algorithmic_hbm"] = {"bytes_per_launch": by / n, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"]}
        return r

    groups = {}
    for name, a in agg.items():
        sym, bound = SYMBOL_OF.get(name, (name, "tensor" if name.startswith(TENSOR_KERNELS) else "hbm"))
        g = groups.setdefault(sym, {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0, "bound": bound, "names": []})
        g["ms"] += a["ms"]; g["bytes"] += a["bytes"]; g["flops"] += a["flops"]; g["n"] += a["n"]; g["names"].append(name)
    sym, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
    roof = record(g["ms"], g["bytes"], g["flops"], g["n"], g["bound"], sym)
    roof["logical_launches"] = sorted(g["names"])
    # DRAM traffic of that symbol per launch, from the committed ncu capture of the same workload (null if not captured)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", f"traffic_{size}_{precision}.json")
    if variant == "flca" and os.path.exists(tpath):
        t = json.load(open(tpath))
        vals = [t[n]["dram_bytes_per_launch"] * agg[n]["n"] for n in g["names"] if n in t and "dram_bytes_per_launch" in t[n]]
        if len(vals) == len(g["names"]):
            traffic = sum(vals) / g["n"]
    roof["traffic"] = traffic
    top_name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    logical = record(top["ms"], top["bytes"], top["flops"], top["n"], "tensor" if top_name.startswith(TENSOR_KERNELS) else "hbm",
                     top_name)
    by_symbol = {k: round(v["ms"] / n_prof, 4) for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}
    return roof, logical, by_symbol


def profile_kernels(model, x_dev, n_prof=3):
    agg = {}
    for _ in range(n_prof):
        _, launches_prof = model.forward_profiled(x_dev)
        for l in launches_prof:
            a = agg.setdefault(l["name"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0})
            a["ms"] += l["ms"]; a["bytes"] += l["bytes"]; a["flops"] += l["flops"]; a["n"] += 1
    return agg


def psnr_db(a, b):
    import math

    import torch

    rng = float(b.max() - b.min())
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * math.log10(rng * rng / mse)


def model_flops(size, variant):
    dim = SIZES[size]
    if variant == "ml":
        return (404.5 * dim * dim + 1005.75 * dim) * 3030272
    return (320.5 * dim * dim + 875.0 * dim) * 3030272


def leg_image_parallel(env, args, size, variant, main):
    """One frame (args.frames) per GPU per step, every rank its own frames, no data-path collective.  `main`: the full
    record of the default workload; else a compact sub-record (BASELINE config 3 at N > 1, the 'sizes' entries at N = 1)."""
    import torch

    import bayer_low_light_image_enhancement_b200 as rf

    B = args.frames
    model = build_model(env, size, variant, args.precision, graphs=not args.no_graph)
    u16_host = synthetic_u16_frames(B, env.rank).pin_memory()
    x_dev = rf.preprocess_u16(u16_host.to(env.dev), **PRE)          # [B,1,H,W] fp32, the frame every leg times
    sampler = ClockSampler(env.local) if main else None
    with torch.no_grad():
        ms_total, launches, host_ms, out = time_resident(env, model, x_dev, args.steps, args.warmup,
                                                         prewarm_s=1.5 if main else 0.5, sampler=sampler)
        clocks = sampler.stop() if (main and env.rank == 0) else None
        out_keep = out.clone() if main else None
        # end to end, the caller's wire formats: uint16 in, uint8 HWC image out
        u8_hosts = [torch.empty(B, H_RAW, W_RAW, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
        ms_e2e, pipe = time_pipeline(env, model, u16_host, u8_hosts, args.steps, preprocess=dict(PRE),
                                     postprocess={"pattern": "RGGB", "auto_rb": True})
        h2d, d2h = pipe.wire_bytes(B, H_RAW, W_RAW)
        assert int(u8_hosts[(args.steps - 1) & 1].max()) > 0          # the result really is on the host
        del pipe
        frames_total = args.steps * B * env.world
        rec = {
            "value": frames_total * MP_FRAME / (ms_total * 1e-3), "unit": "MP/s", "ms_per_step": ms_total / args.steps,
            "e2e": {"value": frames_total * MP_FRAME / (ms_e2e * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                    "api": "FramePipeline(preprocess='u16', postprocess='rgb_u8').submit: uint16 sensor frame H2D, normalise "
                           "(WFB/load_dataset.py:88-89) + forward + clamp/x255/uint8/HWC/channel corrections (test.py:117-120) on "
                           "the device, uint8 image D2H; copies of neighbouring steps overlapped on three streams",
                    "host_binding": env.numa},
            "gpu_launches": launches,
        }
        if not main:
            rec["workload"] = (f"RawFormer-{size} ({variant}), {B} full frame(s) per GPU per step, {args.precision}, "
                               f"image-parallel x{env.world}")
            del model
            torch.cuda.empty_cache()
            return rec
        # the same with fp32 tensors both ways (48.5 MB in, 145 MB out per frame: round 1's e2e definition)
        x_host = x_dev.cpu().pin_memory()
        f32_hosts = [torch.empty(B, 3, H_RAW, W_RAW).pin_memory() for _ in range(2)]
        ms_f32, pipe = time_pipeline(env, model, x_host, f32_hosts, args.steps)
        rec["e2e_fp32_wires"] = {"value": frames_total * MP_FRAME / (ms_f32 * 1e-3), "unit": "MP/s",
                                 "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(f32_hosts[0].numel() * 4),
                                 "ms_per_step": ms_f32 / args.steps}
        del pipe, f32_hosts
        agg = profile_kernels(model, x_dev)
        env.barrier()
    rec.update(host_enqueue_ms_per_step=round(host_ms, 3), clocks=clocks)
    rec["_agg"] = agg
    rec["_out"] = out_keep
    rec["_x_host"] = x_host
    rec["_model"] = model
    return rec


def leg_rowtiled(env, args, size, main):
    """BASELINE config 4: ONE frame per step, cut into row bands over the N ranks (strong scaling).  Every rank holds the
    whole raw frame; halo rows and the per-image reductions cross the GPUs inside kernels of this library (peer-mapped
    memory over NVLink).  N = 1 runs the same band code with a single band."""
    import math

    import torch

    import bayer_low_light_image_enhancement_b200 as rf
    from bayer_low_light_image_enhancement_b200 import _lib

    world, rank, dev = env.world, env.rank, env.dev
    lib = _lib.load()
    variant = args.variant if main else "flca"    # (--row-tiled --variant ml: the multi-level variant row-tiled)
    model = build_model(env, size, variant, "bf16", graphs=False)
    u16_host = synthetic_u16_frames(1, 0).pin_memory()              # the same frame on all ranks
    x_dev = rf.preprocess_u16(u16_host.to(dev), **PRE)
    x_host = x_dev.cpu().pin_memory()
    if world > 1:
        tiled = rf.RowTiledRawFormer.from_process_group(model, H_RAW, W_RAW)
    else:                                         # one band = the whole frame, same band code, no peers
        from bayer_low_light_image_enhancement_b200.rowtiled import _CommRegion

        region = _CommRegion(rf.RowTiledRawFormer.comm_bytes(model, H_RAW, W_RAW, 1), dev)
        tiled = rf.RowTiledRawFormer(model, H_RAW, W_RAW, 0, 1, [region.ptr], own_region=region)
    tiled.check_every = 0                         # (status() synchronises; the bench checks it between its phases)
    band_host = torch.empty(1, tiled.rows, W_RAW, 3, dtype=torch.uint8).pin_memory()
    with torch.no_grad():
        # parity of the decomposition (untimed): bands gathered on rank 0 against the whole-frame forward of the same engine
        full = tiled.gather(tiled(x_dev), dst=0) if world > 1 else tiled(x_dev).clone()
        tiled.status()
        parity = None
        if rank == 0:
            whole = model(x_dev)
            rng = float(whole.max() - whole.min())
            mse = float(((full - whole).double() ** 2).mean())
            parity = {"psnr_vs_whole_frame_db": 99.0 if mse == 0 else 10.0 * math.log10(rng * rng / mse),
                      "max_abs_over_range": float((full - whole).abs().max()) / rng}
            del whole
        del full
        env.barrier()
        t_pre = time.perf_counter()
        n_pre = torch.zeros(1, device=dev)
        while True:                               # every rank must run the same number of forwards: agree on when to stop
            tiled(x_dev)
            n_pre.fill_(1.0 if time.perf_counter() - t_pre < (1.5 if main else 0.5) else 0.0)
            if world > 1:
                env.dist.all_reduce(n_pre, op=env.dist.ReduceOp.MIN)
            if float(n_pre.item()) == 0.0:
                break
        sampler = ClockSampler(env.local) if main else None
        env.barrier()
        if main and rank == 0:
            sampler.start()
        lib.rf_reset_launch_count()
        if not args.no_graph:
            tiled.enable_cuda_graphs()            # this rank's band forward as one graph launch per frame
            env.barrier()
        for _ in range(max(args.warmup, 8)):
            tiled(x_dev)
        env.barrier()
        per_frame = lib.rf_launch_count()         # graph mode: kernels captured once = kernel nodes replayed per frame
        lib.rf_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            band = tiled(x_dev)
        e1.record()
        env.barrier()
        launches = lib.rf_launch_count() if args.no_graph else per_frame * args.steps
        ms_total = env.max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if (main and rank == 0) else None
        tiled.status()
        # end to end in the caller's wire formats: whole uint16 sensor frame host -> every rank (the guidance is replicated),
        # normalised on the device (WFB/load_dataset.py:88-89), forward, this rank's band reduced to uint8 HWC
        # (test.py:117-118) -> host, every step, one stream
        u16_dev = torch.empty_like(u16_host, device=dev)
        for _ in range(2):
            u16_dev.copy_(u16_host, non_blocking=True)
            rf.preprocess_u16(u16_dev, out=x_dev, **PRE)
            band_host.copy_(rf.postprocess_u8(tiled(x_dev)), non_blocking=True)
        env.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            u16_dev.copy_(u16_host, non_blocking=True)
            rf.preprocess_u16(u16_dev, out=x_dev, **PRE)
            band_host.copy_(rf.postprocess_u8(tiled(x_dev)), non_blocking=True)
        f1.record()
        env.barrier()
        ms_e2e = env.max_over_ranks(f0.elapsed_time(f1))
        assert int(band_host.max()) > 0
        # per-kernel times of this rank's band (sync-point kernels include the wait for the peers)
        agg = {}
        n_prof = 3
        for _ in range(n_prof):
            _, lp = tiled.forward_profiled(x_dev)
            for l in lp:
                a = agg.setdefault(l["name"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0})
                a["ms"] += l["ms"]; a["bytes"] += l["bytes"]; a["flops"] += l["flops"]; a["n"] += 1
        tiled.status()
        env.barrier()
    tiled.close()
    del model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    pk = peaks()
    compute = {k: v for k, v in agg.items() if not k.startswith("band_")}
    roof, _, _ = roofline_records(compute, n_prof, pk, size, "bf16", variant)
    roof["scope"] = "rank 0's band"
    rec = {
        "value": args.steps * MP_FRAME / (ms_total * 1e-3), "unit": "MP/s", "ms_per_step": ms_total / args.steps,
        "scaling": "strong",
        "workload": f"RawFormer-{size} ({variant}) forward, ONE SID Sony frame raw {H_RAW}x{W_RAW} per step, row-tiled over "
                    f"{world} GPU(s) (bands of {[r // 16 for _, r in rf.plan_bands(H_RAW, world)]} x 16 rows), random-init weights",
        "parallelism": f"row-tiled x{world}: 4-row halo exchange + all-reduce of the per-image reductions per Conv_Transformer, "
                       "peer-mapped memory over NVLink, no NCCL on the data path",
        "launch": "eager" if args.no_graph else "one CUDA graph per band per frame",
        "e2e": {"value": args.steps * MP_FRAME / (ms_e2e * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": int(u16_host.numel() * 2) * world,
                "d2h_bytes_per_step": 3 * H_RAW * W_RAW, "ms_per_step": ms_e2e / args.steps,
                "api": "preprocess_u16 + RowTiledRawFormer.forward + postprocess_u8 with pinned host buffers: whole uint16 sensor "
                       "frame H2D on every rank, normalise + forward + clamp/x255/uint8/HWC on the device, uint8 band D2H",
                "host_binding": env.numa},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "parity": parity,
        "kernel_ms_per_step": {k: round(v["ms"] / n_prof, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
        "sync_ms_per_step": round(sum(v["ms"] for k, v in agg.items() if k.startswith("band_")) / n_prof, 4),
    }
    return rec


def run_ours(args):
    import torch

    env = Env()
    if args.row_tiled:
        rec = leg_rowtiled(env, args, args.size, main=True)
        env.close()
        if rec is None:
            return
        line = {"metric": METRIC, "value": rec.pop("value"), "unit": "MP/s", "n_gpus": env.world, "steps": args.steps,
                "warmup": max(args.warmup, 8), "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": rec.pop("workload"), "precision": "bf16", "parallelism": rec.pop("parallelism"),
                           "l2": "per-step working set >> 126 MB L2, no explicit flush", "launch": rec.pop("launch")}}
        rec.pop("unit", None); rec.pop("scaling", None)
        line.update(rec)
        print(json.dumps(line), flush=True)
        return

    rec = leg_image_parallel(env, args, args.size, args.variant, main=True)
    agg, out_gpu, x_host, model = rec.pop("_agg"), rec.pop("_out"), rec.pop("_x_host"), rec.pop("_model")
    extra = {}
    if not args.no_extra:
        with torch.no_grad():
            if env.world == 1:
                # the other BASELINE model sizes / the multi-level variant on the same frame (device-resident ms per frame)
                sizes = {}
                del model
                torch.cuda.empty_cache()
                x_dev = x_host.to(env.dev)
                for name, (sz, var) in (("B", ("B", "flca")), ("L", ("L", "flca")), ("ml_S", ("S", "ml"))):
                    if (sz, var) == (args.size, args.variant):
                        continue
                    m = build_model(env, sz, var, args.precision, graphs=not args.no_graph)
                    ms, _, _, _ = time_resident(env, m, x_dev, max(3, args.steps // 2), 3, prewarm_s=0.3, load_s=0.2)
                    n = max(3, args.steps // 2)
                    sizes[name] = {"ms_per_frame": round(ms / n, 4), "value": round(MP_FRAME / (ms / n * 1e-3), 1), "unit": "MP/s",
                                   "model": f"RawFormer-{sz} ({var})"}
                    del m
                    torch.cuda.empty_cache()
                extra["sizes"] = sizes
            else:
                del model
                torch.cuda.empty_cache()
                if args.size != "B":
                    extra["config3"] = leg_image_parallel(env, args, "B", "flca", main=False)
                torch.cuda.empty_cache()
                try:
                    extra["config4"] = leg_rowtiled(env, args, "L", main=False)
                except Exception as e:  # noqa: BLE001  (the default S line must survive a failure of the extra legs)
                    extra["config4"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if env.rank != 0:
        env.close()
        return

    pk = peaks()
    n_prof = 3
    roof, logical, by_symbol = roofline_records(agg, n_prof, pk, args.size, args.precision, args.variant)
    breakdown = sorted(((k, v["ms"] / n_prof) for k, v in agg.items()), key=lambda kv: -kv[1])
    frames_total = args.steps * args.frames * env.world
    cpu = parity = None
    if env.world == 1 and not args.no_cpu:
        cpu, ref = time_cpu_port(args.size, args.variant, x_host)
        p = psnr_db(out_gpu.cpu(), ref)
        parity = {"psnr_vs_reference_cpu_forward_db": round(p, 2),
                  "max_abs": float((out_gpu.cpu() - ref).abs().max()), "ref_range": float(ref.max() - ref.min()),
                  "what": "output of the timed frame (bf16 engine) against the reference's fp32 CPU forward of the same frame and weights"}
    eager = None
    if env.world == 1 and args.torch_eager:
        torch.cuda.empty_cache()
        eager = time_torch_eager_b200(args.size, args.variant, env.dev)
    flops = model_flops(args.size, args.variant)
    bytes_per_frame = sum(v["bytes"] for v in agg.values()) / n_prof
    line = {
        "metric": METRIC, "value": rec["value"], "unit": "MP/s", "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, env.world), "precision": args.precision,
                   "parallelism": f"image-parallel x{env.world}" if env.world > 1 else "single GPU",
                   "input": "synthetic uint16 sensor frames (black level 512) normalised like WFB/load_dataset.py:88-89 at ratio 100",
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2, no explicit flush",
                   "launch": "eager" if args.no_graph else "one CUDA graph per frame (model.enable_cuda_graphs)"},
        "e2e": rec["e2e"], "e2e_fp32_wires": rec["e2e_fp32_wires"],
        "gpu_launches": rec["gpu_launches"], "host_enqueue_ms_per_step": rec["host_enqueue_ms_per_step"],
        "clocks": rec["clocks"], "roofline": roof, "roofline_logical": logical,
        "kernel_ms_per_step": {k: round(v, 4) for k, v in breakdown}, "symbol_ms_per_step": by_symbol,
        "model_roofline": {"flops_per_frame": flops, "achieved_tflops": flops * frames_total / (rec["ms_per_step"] * args.steps * 1e-3) / 1e12,
                           "algorithmic_bytes_per_frame_sum_of_launches": bytes_per_frame},
    }
    if parity is not None:
        line["parity_db"] = parity["psnr_vs_reference_cpu_forward_db"]
        line["parity"] = parity
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if eager is not None:
        line["eager_b200_baseline"] = eager
    line.update(extra)
    print(json.dumps(line), flush=True)
    env.close()


def run_ops(args):
    """--op dwt: north-star subsystem (1), the wavelet / shuffle operators as stand-alone HBM streaming kernels at the shape the
    README quotes ([1,32,1424,2128] fp32 <-> [1,128,712,1064]): achieved GB/s = algorithmic bytes (read once + write once)
    / CUDA-event time, against the measured copy bandwidth.  388 MB per tensor >> 126 MB L2: no flush needed."""
    import torch

    import bayer_low_light_image_enhancement_b200 as rf

    env = Env()
    dev = env.dev
    pk = peaks()
    x = torch.randn(1, 32, 1424, 2128, device=dev)
    sub = torch.randn(1, 128, 712, 1064, device=dev)
    raw = torch.rand(1, 1, 2848, 4256, device=dev)
    kh = [[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]
    dwt, idwt = rf.CustomDWT().to(dev), rf.CustomIDWT().to(dev)
    dwt_h, idwt_h = rf.CustomDWT(kernel=kh).to(dev), rf.CustomIDWT(kernel=kh).to(dev)
    haar, ps = rf.HaarDWT().to(dev), rf.PixelShuffle(2)
    ops = [
        ("CustomDWT  [1,32,1424,2128] -> [1,128,712,1064] (README.md:111-117)", lambda: dwt(x), 2 * x.numel() * 4),
        ("CustomIDWT [1,128,712,1064] -> [1,32,1424,2128] (README.md:139-144)", lambda: idwt(sub), 2 * sub.numel() * 4),
        ("CustomDWT, Haar matrix", lambda: dwt_h(x), 2 * x.numel() * 4),
        ("CustomIDWT, Haar matrix", lambda: idwt_h(sub), 2 * sub.numel() * 4),
        ("dwt_init   [1,32,1424,2128] -> [4,32,712,1064] (WFB/blocks.py:102-115)", lambda: rf.dwt_init(x), 2 * x.numel() * 4),
        ("iwt_init   [4,32,712,1064] -> [1,32,1424,2128] (WFB/blocks.py:119-136)",
         lambda: rf.iwt_init(sub.view(4, 32, 712, 1064)), 2 * sub.numel() * 4),
        ("HaarDWT    [1,32,1424,2128] -> 4 x [1,32,712,1064] (FLCA_RF.py:56-73)", lambda: haar(x), 2 * x.numel() * 4),
        ("downshuffle r=2 [1,32,1424,2128] (FLCA_RF.py:18-33)", lambda: rf.downshuffle(x, 2), 2 * x.numel() * 4),
        ("PixelShuffle(2) [1,128,712,1064] (FLCA_RF.py:328)", lambda: ps(sub), 2 * sub.numel() * 4),
        ("downshuffle r=2, raw frame [1,1,2848,4256]", lambda: rf.downshuffle(raw, 2), 2 * raw.numel() * 4),
    ]
    out = []
    with torch.no_grad():
        for name, fn, nbytes in ops:
            for _ in range(max(args.warmup, 3)):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            gbs = nbytes / (ms * 1e-3) / 1e9
            out.append({"op": name, "ms": round(ms, 4), "bytes": nbytes, "achieved_gbs": round(gbs, 1), "frac": round(gbs / pk["hbm_gbs"], 3)})
    line = {"metric": "HBM GB/s of the wavelet / shuffle operators (algorithmic bytes / CUDA-event time, includes the torch.empty of "
                      "the output tensor)", "unit": "GB/s", "peak": pk["hbm_gbs"], "peak_source": pk["src"], "steps": args.steps,
            "ops": out}
    print(json.dumps(line), flush=True)
    env.close()


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the extra legs (N = 1: other model sizes; N > 1: BASELINE configs 3 and 4 sub-records)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of one CUDA graph per frame")
    ap.add_argument("--torch-eager", action="store_true",
                    help="also time the reference's operator sequence run eagerly by PyTorch on this B200 (baseline leg)")
    ap.add_argument("--row-tiled", action="store_true",
                    help="BASELINE config 4: one frame per step cut into row bands over the ranks (strong scaling)")
    ap.add_argument("--op", default=None, choices=["dwt"],
                    help="time the stand-alone wavelet / shuffle operators instead of the model (north-star subsystem 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.op == "dwt":
        run_ops(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
