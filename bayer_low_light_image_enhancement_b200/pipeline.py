"""Frame streaming around the RawFormer forward: the caller side of ``test.py:100-118`` (DataLoader frame in pinned
host memory -> ``model(inp)`` -> ``.cpu()``), with the host<->device copies of consecutive frames overlapped with the
forward on three CUDA streams.  PCIe moves 48 MB in and 145 MB out per SID Sony frame; run back to back on one stream
that is a third of the step, overlapped it is hidden behind the forward."""
from __future__ import annotations

import torch


class FramePipeline:
    """``submit(x_host, out_host)`` enqueues H2D copy -> forward -> D2H copy of one batch of frames and returns at once;
    ``flush()`` waits for everything submitted.  ``x_host`` [B,1,H,W] and ``out_host`` [B,3,H,W] should be pinned fp32
    tensors; ``out_host`` is valid after ``flush()`` (or after ``wait(ticket)``).  Frames are processed in order;
    ``depth`` device-side slots let copy-in of frame i+1 and copy-out of frame i-1 run under the forward of frame i.
    With ``graphs`` (default) the model replays one CUDA graph per slot instead of launching ~130 kernels per frame."""

    def __init__(self, model, depth: int = 2, graphs: bool = True):
        p = next(model.parameters(), None)
        if p is None or not p.is_cuda:
            raise RuntimeError("FramePipeline needs a model on a CUDA device (there is no CPU path)")
        self.model = model
        self.dev = p.device
        self.depth = max(1, int(depth))
        # copies on high-priority streams (a different stream pool than the compute stream): with the default of 8 hardware
        # connections, three same-priority pool streams can alias one connection and the copies then serialise with the
        # forward (seen as a bimodal 8 ms / 13 ms per frame); bench.py also raises CUDA_DEVICE_MAX_CONNECTIONS
        self.s_in = torch.cuda.Stream(self.dev, priority=-1)
        self.s_cmp = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev, priority=-1)
        self.slots = [dict(x=None, out=None, ev_in=torch.cuda.Event(), ev_cmp=torch.cuda.Event(),
                           ev_out=torch.cuda.Event(), used=False) for _ in range(self.depth)]
        self.n = 0
        if graphs and hasattr(model, "enable_cuda_graphs"):
            model.enable_cuda_graphs(True, max_graphs=max(8, 2 * self.depth))   # one graph per input slot

    def start_after(self, event: torch.cuda.Event):
        """Make the pipeline's streams wait for ``event`` (e.g. a timing event recorded on the current stream)."""
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(event)

    @torch.no_grad()
    def submit(self, x_host: torch.Tensor, out_host: torch.Tensor):
        slot = self.slots[self.n % self.depth]
        self.n += 1
        if slot["x"] is None or slot["x"].shape != x_host.shape:
            slot["x"] = torch.empty(x_host.shape, dtype=torch.float32, device=self.dev)
        if slot["used"]:
            self.s_in.wait_event(slot["ev_cmp"])     # the forward that read this slot's input has finished
        with torch.cuda.stream(self.s_in):
            slot["x"].copy_(x_host, non_blocking=True)
            slot["ev_in"].record(self.s_in)
        self.s_cmp.wait_event(slot["ev_in"])
        if slot["used"]:
            self.s_cmp.wait_event(slot["ev_out"])    # the previous result of this slot has left the device
        with torch.cuda.stream(self.s_cmp):
            out = self.model(slot["x"])
            slot["ev_cmp"].record(self.s_cmp)
        self.s_out.wait_event(slot["ev_cmp"])
        with torch.cuda.stream(self.s_out):
            out_host.copy_(out, non_blocking=True)
            slot["ev_out"].record(self.s_out)
        out.record_stream(self.s_out)
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
