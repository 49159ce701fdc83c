"""Frame streaming around the RawFormer forward: the caller side of ``test.py:100-120`` (DataLoader frame in pinned
host memory -> ``model(inp)`` -> ``clamp(0,1) -> .cpu() -> *255 -> uint8`` -> channel corrections), with the host<->device
copies of consecutive frames overlapped with the forward on three CUDA streams.

Two wire formats per direction (one SID Sony frame, raw 2848x4256):

* input  ``fp32``  [B,1,H,W] float32, 48.5 MB (what the reference's DataLoader yields, ``WFB/load_dataset.py:91``), or
         ``u16``   [B,H,W] uint16 sensor values, 24.2 MB: the normalisation of ``WFB/load_dataset.py:88-89`` (black level,
         white level, exposure ratio; optional clamp of ``correctdataloader.py:103``) runs on the device
         (``rf_preprocess_u16``) straight into the model's input buffer;
* output ``fp32``  [B,3,H,W] float32, 145 MB (``pred``), or
         ``rgb_u8`` [B,H,W,3] uint8, 36.4 MB: what ``test.py:117-120`` reduces ``pred`` to immediately -- clamp, x255,
         truncation to uint8, HWC, ``correct_bayer_channels(pattern)`` and the data-dependent ``auto_correct_rb`` -- done
         on the device (``rf_postprocess_rgb_u8``) before the copy.

With ``u16`` in and ``rgb_u8`` out a frame costs 60.6 MB of PCIe traffic instead of 194 MB, which is what bounds the
8-GPU image-parallel run on one host (round 1: 0.52 end-to-end scaling efficiency at 8 GPUs with fp32 wires).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .extras import _perm


class FramePipeline:
    """``submit(x_host, out_host)`` enqueues H2D copy -> [normalise] -> forward -> [uint8 conversion] -> D2H copy of one
    batch of frames and returns at once; ``flush()`` waits for everything submitted.  Frames are processed in order;
    ``depth`` device-side slots let copy-in of frame i+1 and copy-out of frame i-1 run under the forward of frame i.
    With ``graphs`` (default) the model replays one CUDA graph per slot instead of launching its kernels one by one.

    ``preprocess``: None -> ``x_host`` is float32 [B,1,H,W]; ``"u16"`` or a dict ``{"black": 512, "white": 16383, "ratio":
    100 (a number, or one per submit via ``submit(..., ratio=)``), "clamp": True}`` -> ``x_host`` is uint16 [B,H,W].
    ``postprocess``: None -> ``out_host`` is float32 [B,3,H,W]; ``"rgb_u8"`` or a dict ``{"pattern": "RGGB", "auto_rb":
    True}`` -> ``out_host`` is uint8 [B,H,W,3].  Host tensors should be pinned; ``out_host`` is valid after ``flush()``
    (or after ``wait(ticket)``)."""

    def __init__(self, model, depth: int = 2, graphs: bool = True, preprocess=None, postprocess=None):
        p = next(model.parameters(), None)
        if p is None or not p.is_cuda:
            raise RuntimeError("FramePipeline needs a model on a CUDA device (there is no CPU path)")
        self.model = model
        self.dev = p.device
        self.depth = max(1, int(depth))
        self.pre = self._pre_cfg(preprocess)
        self.post = self._post_cfg(postprocess)
        _lib.init_device(self.dev)
        # copies on high-priority streams (a different stream pool than the compute stream): with the default of 8 hardware
        # connections, three same-priority pool streams can alias one connection and the copies then serialise with the
        # forward (seen as a bimodal 8 ms / 13 ms per frame); bench.py also raises CUDA_DEVICE_MAX_CONNECTIONS
        self.s_in = torch.cuda.Stream(self.dev, priority=-1)
        self.s_cmp = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev, priority=-1)
        self.slots = [dict(x=None, raw=None, u8=None, sums=None, out=None, ev_in=torch.cuda.Event(), ev_cmp=torch.cuda.Event(),
                           ev_out=torch.cuda.Event(), used=False) for _ in range(self.depth)]
        self.n = 0
        if graphs and hasattr(model, "enable_cuda_graphs"):
            model.enable_cuda_graphs(True, max_graphs=max(8, 2 * self.depth))   # one graph per input slot

    @staticmethod
    def _pre_cfg(cfg):
        if cfg is None:
            return None
        if cfg == "u16":
            cfg = {}
        if not isinstance(cfg, dict):
            raise ValueError("preprocess must be None, 'u16' or a dict")
        unknown = set(cfg) - {"kind", "black", "white", "ratio", "clamp"}
        if unknown or cfg.get("kind", "u16") != "u16":
            raise ValueError(f"unsupported preprocess configuration {cfg!r}")
        return dict(black=float(cfg.get("black", 512.0)), white=float(cfg.get("white", 16383.0)),
                    ratio=float(cfg.get("ratio", 100.0)), clamp=bool(cfg.get("clamp", True)))

    @staticmethod
    def _post_cfg(cfg):
        if cfg is None:
            return None
        if cfg == "rgb_u8":
            cfg = {}
        if not isinstance(cfg, dict):
            raise ValueError("postprocess must be None, 'rgb_u8' or a dict")
        unknown = set(cfg) - {"kind", "pattern", "auto_rb"}
        if unknown or cfg.get("kind", "rgb_u8") != "rgb_u8":
            raise ValueError(f"unsupported postprocess configuration {cfg!r}")
        return dict(pattern=str(cfg.get("pattern", "RGGB")), auto_rb=bool(cfg.get("auto_rb", True)))

    def wire_bytes(self, batch, H, W):
        """(host->device, device->host) bytes per submit of ``batch`` frames of H x W raw pixels."""
        return batch * H * W * (2 if self.pre else 4), batch * 3 * H * W * (1 if self.post else 4)

    def start_after(self, event: torch.cuda.Event):
        """Make the pipeline's streams wait for ``event`` (e.g. a timing event recorded on the current stream)."""
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(event)

    def _check_host(self, x_host, out_host):
        if self.pre is None:
            if x_host.dim() != 4 or x_host.shape[1] != 1 or x_host.dtype != torch.float32:
                raise ValueError("x_host must be float32 [B,1,H,W] (or construct the pipeline with preprocess='u16')")
            b, _, h, w = x_host.shape
        else:
            if x_host.dim() == 4 and x_host.shape[1] == 1:
                x_host = x_host[:, 0]
            if x_host.dim() != 3 or x_host.dtype not in (torch.uint16, torch.int16):
                raise ValueError("x_host must be a 16-bit integer tensor [B,H,W] with preprocess='u16'")
            b, h, w = x_host.shape
        want = (b, h, w, 3) if self.post else (b, 3, h, w)
        if tuple(out_host.shape) != want or out_host.dtype != (torch.uint8 if self.post else torch.float32):
            raise ValueError(f"out_host must be {'uint8' if self.post else 'float32'} {list(want)}")
        if not x_host.is_contiguous() or not out_host.is_contiguous():
            raise ValueError("host tensors must be contiguous")
        return x_host, b, h, w

    @torch.no_grad()
    def submit(self, x_host: torch.Tensor, out_host: torch.Tensor, ratio=None):
        x_host, b, h, w = self._check_host(x_host, out_host)
        slot = self.slots[self.n % self.depth]
        self.n += 1
        lib = _lib.load()
        if slot["x"] is None or tuple(slot["x"].shape) != (b, 1, h, w):
            # device buffers of the slot, allocated on the streams that use them first (caching-allocator ownership)
            with torch.cuda.stream(self.s_cmp):
                if slot["used"]:
                    self.s_cmp.wait_event(slot["ev_out"])
                slot["x"] = torch.empty(b, 1, h, w, dtype=torch.float32, device=self.dev)
                slot["u8"] = torch.empty(b, h, w, 3, dtype=torch.uint8, device=self.dev) if self.post else None
                slot["sums"] = torch.empty(2 * b, dtype=torch.int64, device=self.dev) if self.post else None
            with torch.cuda.stream(self.s_in):
                slot["raw"] = torch.empty(b, h, w, dtype=x_host.dtype, device=self.dev) if self.pre else None
            self.s_in.wait_stream(self.s_cmp)
            slot["x"].record_stream(self.s_in)
        if slot["used"]:
            self.s_in.wait_event(slot["ev_cmp"])     # the forward that read this slot's input has finished
        with torch.cuda.stream(self.s_in):
            (slot["raw"] if self.pre else slot["x"]).copy_(x_host, non_blocking=True)
            slot["ev_in"].record(self.s_in)
        self.s_cmp.wait_event(slot["ev_in"])
        if slot["used"]:
            self.s_cmp.wait_event(slot["ev_out"])    # the previous result of this slot has left the device
        with torch.cuda.stream(self.s_cmp):
            if self.pre:
                c = self.pre
                check(lib.rf_preprocess_u16(ptr(slot["raw"]), ptr(slot["x"]), c["black"], c["white"],
                                            float(c["ratio"] if ratio is None else ratio), int(c["clamp"]), b, h, w,
                                            stream_ptr(self.dev)), "rf_preprocess_u16")
            out = self.model(slot["x"])
            if self.post:
                c = self.post
                if not out.is_contiguous() or out.dtype != torch.float32:
                    out = out.float().contiguous()
                check(lib.rf_postprocess_rgb_u8(ptr(out), ptr(slot["u8"]), _perm(c["pattern"]), int(c["auto_rb"]), b, h, w,
                                                ptr(slot["sums"]), slot["sums"].numel() * 8, stream_ptr(self.dev)),
                      "rf_postprocess_rgb_u8")
                result = slot["u8"]
            else:
                result = out
            slot["ev_cmp"].record(self.s_cmp)
        self.s_out.wait_event(slot["ev_cmp"])
        with torch.cuda.stream(self.s_out):
            out_host.copy_(result, non_blocking=True)
            slot["ev_out"].record(self.s_out)
        result.record_stream(self.s_out)
        slot["out"] = out
        slot["used"] = True
        return slot["ev_out"]

    def wait(self, ticket: torch.cuda.Event):
        ticket.synchronize()

    def finish_event(self) -> torch.cuda.Event:
        """An event on the CURRENT stream that completes after everything submitted so far."""
        cur = torch.cuda.current_stream(self.dev)
        for slot in self.slots:
            if slot["used"]:
                cur.wait_event(slot["ev_out"])
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(cur)
        return ev

    def flush(self):
        self.s_out.synchronize()
        self.s_cmp.synchronize()
