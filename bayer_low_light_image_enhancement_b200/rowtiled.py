"""Row-tiled single-frame forward (BASELINE config 4): ONE frame is cut into bands of whole rows, one band
per GPU.  Every rank gets the whole raw frame (the 1-channel guidance of ``FLCA_RF.py:87-97,140-148`` is replicated),
runs the U-Net on its band and, per ``Conv_Transformer``, exchanges 4 halo rows of the block input with its band
neighbours and all-reduces the block's three per-image reductions (squeeze-excite channel sums ``FLCA_RF.py:160``,
``|q|^2,|k|^2`` and the per-head Gram ``FLCA_RF.py:228-230``).  Both steps are kernels of ``librawformer_b200.so`` that
write the peers' memory directly over NVLink (CUDA-IPC mapped "comm regions"); ``torch.distributed`` is used only to
hand the IPC handles round once and, optionally, to gather the bands on one rank.

    tiled = RowTiledRawFormer.from_process_group(model, H, W)     # one process per GPU, after init_process_group
    band = tiled(raw)                                            # raw [1,1,H,W] -> this rank's rows [1,3,rows,W]
    full = tiled.gather(band, dst=0)

``LocalBands`` runs the same ranks as concurrent streams of ONE GPU (test vehicle: same kernels, same protocol).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr

RF_BAND_MAX_RANKS = 8
RF_BAND_HALO = 4
RF_IPC_HANDLE_BYTES = 64
MIN_BAND_ROWS = 16 * RF_BAND_HALO


def plan_bands(H: int, nranks: int):
    """[(row0, rows)] per rank: contiguous bands in units of 16 raw rows (every U-Net stage then has whole rows;
    SURVEY 8d: 2848 rows = 178 units -> 89+89, 44+45+44+45, 6 x 22 + 2 x 23).  A rank works on its band PLUS its halo
    rows (half a unit per neighbour), so the units left over after an even split go to the ranks whose band image is
    smallest: the two border ranks first (178 units over 4 ranks -> 45, 44, 44, 45)."""
    if H <= 0 or H % 16:
        raise ValueError(f"frame height {H} must be a positive multiple of 16")
    if not 1 <= nranks <= RF_BAND_MAX_RANKS:
        raise ValueError(f"nranks must be in 1..{RF_BAND_MAX_RANKS}")
    units = H // 16
    if units < nranks * RF_BAND_HALO:
        raise ValueError(f"{H} rows are too few for {nranks} bands of at least {MIN_BAND_ROWS} rows")
    size = [units // nranks] * nranks
    halo = [(1 if r > 0 else 0) + (1 if r < nranks - 1 else 0) for r in range(nranks)]     # in half units
    for _ in range(units % nranks):
        r = min(range(nranks), key=lambda i: (2 * size[i] + halo[i], min(i, nranks - 1 - i), i))
        size[r] += 1
    row0 = [16 * sum(size[:r]) for r in range(nranks)]
    return [(row0[r], 16 * size[r]) for r in range(nranks)]


class _CommRegion:
    """A comm region of this process' current device (flags + mailboxes), exportable to the peers."""

    def __init__(self, nbytes: int, device: torch.device):
        lib = _lib.load()
        self.device = device
        self.nbytes = int(nbytes)
        p = C.c_void_p(0)
        h = (C.c_ubyte * RF_IPC_HANDLE_BYTES)()
        with torch.cuda.device(device):
            check(lib.rf_band_comm_alloc(self.nbytes, C.byref(p), h), "rf_band_comm_alloc")
        self.ptr = p.value
        self.handle = bytes(h)

    def free(self):
        if self.ptr:
            with torch.cuda.device(self.device):
                _lib.load().rf_band_comm_free(C.c_void_p(self.ptr))
            self.ptr = 0


class RowTiledRawFormer:
    """This rank's share of the row-tiled forward.  ``comm_ptrs[r]`` = rank r's comm region as mapped in this process."""

    def __init__(self, model, H: int, W: int, rank: int, nranks: int, comm_ptrs, own_region=None, group=None,
                 opened=()):
        if getattr(model, "variant", None) not in (_lib.RF_VARIANT_FLCA, _lib.RF_VARIANT_ML):
            raise NotImplementedError("row tiling implements FLCA_RF.py::RawFormer (config 4) and ML_RF.py::RawFormer")
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("RowTiledRawFormer needs a model on a CUDA device (there is no CPU path)")
        if model._dtype() != _lib.RF_BF16:
            raise NotImplementedError("row tiling runs the bf16 tensor-core engine (precision='bf16')")
        self.model, self.device = model, p.device
        self.H, self.W, self.rank, self.nranks = int(H), int(W), int(rank), int(nranks)
        self.row0, self.rows = plan_bands(self.H, self.nranks)[self.rank]
        self.group, self._own, self._opened = group, own_region, list(opened)
        self.epoch = 0
        self._graphs = None
        self.check_every = 64          # forward() reads the sticky error word every this many frames (synchronises)
        lib = _lib.load()
        self._band = _lib.Band()
        self._band.rank, self._band.nranks = self.rank, self.nranks
        self._band.row0, self._band.rows = self.row0, self.rows
        for r in range(self.nranks):
            self._band.comm[r] = comm_ptrs[r]
        orows, r0 = C.c_int(0), C.c_int(0)
        check(lib.rf_band_out_rows(C.byref(self._band), C.byref(orows), C.byref(r0)), "rf_band_out_rows")
        self._out_rows, self._out_r0 = orows.value, r0.value
        nws = lib.rf_rawformer_band_workspace_bytes(model.dim, _lib.RF_BF16, model.variant, self.H, self.W,
                                                    C.byref(self._band))
        if nws == 0:
            raise ValueError(f"unsupported row-tiled configuration: dim {model.dim}, frame {H}x{W}, {nranks} bands")
        self._ws = torch.zeros(nws, dtype=torch.uint8, device=self.device)   # own: bands of one GPU run concurrently
        self.rehearse()

    @torch.no_grad()
    def rehearse(self):
        """One forward with epoch 0 (no rank signals or waits) on a zero frame: loads every kernel of this rank's plan.
        The first launch of a kernel may wait for the device to go idle (lazy module loading); inside a real frame that
        would stall behind a sync-point kernel that is itself waiting for a peer."""
        raw = torch.zeros(1, 1, self.H, self.W, dtype=torch.float32, device=self.device)
        self._launch(raw, 0)
        torch.cuda.current_stream(self.device).synchronize()

    # -- construction ------------------------------------------------------------------------------------------------
    @staticmethod
    def comm_bytes(model, H, W, nranks):
        n = _lib.load().rf_band_comm_bytes(model.dim, _lib.RF_BF16, model.variant, int(H), int(W), int(nranks))
        if n == 0:
            raise ValueError(f"unsupported row-tiled configuration: dim {model.dim}, frame {H}x{W}, {nranks} bands")
        return n

    @classmethod
    def from_process_group(cls, model, H, W, group=None):
        """One process per GPU: allocate this rank's comm region, swap CUDA-IPC handles, map the peers' regions."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = next(model.parameters()).device
        _lib.init_device(dev)
        own = _CommRegion(cls.comm_bytes(model, H, W, world), dev)
        handles = [None] * world
        dist.all_gather_object(handles, own.handle, group=group)
        lib = _lib.load()
        ptrs, opened = [], []
        with torch.cuda.device(dev):
            for r in range(world):
                if r == rank:
                    ptrs.append(own.ptr)
                    continue
                p = C.c_void_p(0)
                h = (C.c_ubyte * RF_IPC_HANDLE_BYTES).from_buffer_copy(handles[r])
                check(lib.rf_band_comm_open(h, C.byref(p)), f"rf_band_comm_open(rank {r})")
                ptrs.append(p.value)
                opened.append(p.value)
        dist.barrier(group=group)
        tiled = cls(model, H, W, rank, world, ptrs, own_region=own, group=group, opened=opened)
        # the constructor allocates GBs, loads every kernel (rehearsal) and packs the weights: without this barrier the
        # rank skew at the first sync point of the first real frame is unbounded (the waits time out after ~6 s)
        torch.cuda.current_stream(tiled.device).synchronize()
        dist.barrier(group=group)
        return tiled

    def close(self):
        """Unmap the peers' comm regions and free this rank's.  The peers write into this rank's region until their last
        forward has finished: synchronise and barrier the group first."""
        lib = _lib.load()
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.rf_band_comm_close(C.c_void_p(p))
        self._opened = []
        if self._own is not None:
            self._own.free()
            self._own = None

    # -- forward ---------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, raw: torch.Tensor) -> torch.Tensor:
        """raw [1,1,H,W] (the WHOLE frame, on this rank's GPU) -> this rank's rows of the result, [1,3,rows,W] (a view
        of the band image the kernels write).  Every rank must call it the same number of times."""
        m = self.model
        raw = m._check_input(raw)
        if tuple(raw.shape) != (1, 1, self.H, self.W):
            raise ValueError(f"expected the whole frame [1,1,{self.H},{self.W}], got {tuple(raw.shape)}")
        self.epoch += 1
        out = self._replay(raw) if self._graphs is not None else self._launch(raw, self.epoch)
        if self.check_every and self.epoch % self.check_every == 0:
            self.status()          # a peer that never arrived makes every later frame wrong: fail loudly, not silently
        return out

    __call__ = forward

    def reset(self):
        """Recover from a failed frame (``status()`` raised on some rank, or a rank raised mid-frame): every rank calls
        this.  Waits for the local stream, meets the group, clears the local region's error word / frame counter / arrival
        counters, meets the group again and restarts the frame count.  Captured graphs stay valid (the frame counter they
        compare against lives in the region)."""
        torch.cuda.current_stream(self.device).synchronize()
        if self.group is not None or self._opened:
            import torch.distributed as dist

            dist.barrier(group=self.group)
        check(_lib.load().rf_band_comm_reset(C.c_void_p(self._band.comm[self.rank]), _lib.stream_ptr(self.device)),
              "rf_band_comm_reset")
        if self.group is not None or self._opened:
            import torch.distributed as dist

            dist.barrier(group=self.group)
        self.epoch = 0

    def enable_cuda_graphs(self, on=True):
        """Replay this rank's forward as ONE CUDA graph per input buffer (the frame counter the sync points compare
        against lives in the comm region and is advanced on the device, so the launch sequence is the same every frame).
        The result is the graph's own output tensor: it is overwritten by the next call with the same input buffer."""
        self._graphs = {} if on else None
        return self

    def capture(self, raw):
        """Capture (without running) the graph of this rank's forward for the input buffer ``raw``."""
        blob = self.model.packed_weights(raw.device, _lib.RF_BF16)      # pack BEFORE the capture (it synchronises)
        key = (raw.data_ptr(), tuple(raw.shape), blob.data_ptr())
        hit = self._graphs.get(key)
        if hit is None:
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))              # oldest entry only
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._launch(raw, 1)
            hit = (g, out, raw, blob)                                   # keeps input storage and packed weights alive
            self._graphs[key] = hit
        return hit

    def _replay(self, raw):
        hit = self.capture(raw)
        hit[0].replay()
        return hit[1]

    def _launch(self, raw, epoch):
        m, lib = self.model, _lib.load()
        blob = m.packed_weights(raw.device, _lib.RF_BF16)
        out = torch.empty(1, 3, self._out_rows, self.W, dtype=torch.float32, device=raw.device)
        self._band.epoch = epoch
        check(lib.rf_rawformer_forward_band(ptr(blob), m.dim, _lib.RF_BF16, m.variant, ptr(raw), ptr(out), self.H, self.W,
                                            C.byref(self._band), ptr(self._ws), self._ws.numel(),
                                            _lib.stream_ptr(raw.device)), "rf_rawformer_forward_band")
        return out[:, :, self._out_r0:self._out_r0 + self.rows]

    @torch.no_grad()
    def forward_profiled(self, raw, cap=4096):
        """One forward with a CUDA-event pair around every launch (synchronises; every rank must call it).  Returns
        (band, launches) like ``RawFormer.forward_profiled``; band_halo / band_allreduce times include the wait for peers."""
        m, lib = self.model, _lib.load()
        raw = m._check_input(raw)
        blob = m.packed_weights(raw.device, _lib.RF_BF16)
        out = torch.empty(1, 3, self._out_rows, self.W, dtype=torch.float32, device=raw.device)
        self.epoch += 1
        self._band.epoch = self.epoch
        ms, ids, n = (C.c_float * cap)(), (C.c_int * cap)(), C.c_int(0)
        check(lib.rf_rawformer_forward_band_profiled(ptr(blob), m.dim, _lib.RF_BF16, m.variant, ptr(raw), ptr(out), self.H,
                                                     self.W, C.byref(self._band), ptr(self._ws), self._ws.numel(),
                                                     _lib.stream_ptr(raw.device), ms, ids, cap, C.byref(n)),
              "rf_rawformer_forward_band_profiled")
        launches, by, fl = [], C.c_double(0), C.c_double(0)
        for i in range(min(n.value, cap)):
            lib.rf_profiled_launch_info(i, C.byref(by), C.byref(fl))
            launches.append({"name": lib.rf_kernel_name(ids[i]).decode(), "ms": float(ms[i]), "bytes": by.value,
                             "flops": fl.value})
        return out[:, :, self._out_r0:self._out_r0 + self.rows], launches

    def status(self):
        """Synchronise and raise if a cross-GPU wait of this rank timed out (a peer never reached the sync point)."""
        err = C.c_int(0)
        check(_lib.load().rf_band_comm_status(C.c_void_p(self._band.comm[self.rank]), C.byref(err),
                                              _lib.stream_ptr(self.device)), "rf_band_comm_status")
        if err.value:
            raise RuntimeError(f"row-tiled forward: rank {self.rank} timed out at sync point {err.value - 1}")

    def gather(self, band: torch.Tensor, dst: int = 0):
        """All bands on rank ``dst`` as one [1,3,H,W] frame (NCCL point-to-point: the bands may differ in height);
        None on the other ranks."""
        import torch.distributed as dist

        self.status()                      # never hand out bands of a frame whose sync points timed out
        band = band.contiguous()
        me = dist.get_rank(self.group)
        if me != dst:
            ops = [dist.P2POp(dist.isend, band, dst, self.group)]
            parts = None
        else:
            parts = [band if r == me else torch.empty(1, 3, rows, self.W, dtype=band.dtype, device=band.device)
                     for r, (_, rows) in enumerate(plan_bands(self.H, self.nranks))]
            ops = [dist.P2POp(dist.irecv, parts[r], r, self.group) for r in range(self.nranks) if r != me]
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return torch.cat(parts, dim=2) if parts is not None else None


class LocalBands:
    """All bands of a frame as concurrent streams of ONE GPU: the same kernels and the same flag/mailbox protocol as
    the multi-GPU run, with the "peer" regions in local memory.  For parity tests of the decomposition."""

    def __init__(self, model, H, W, nranks, graphs=False):
        import os

        dev = next(model.parameters()).device
        _lib.init_device(dev)
        # the bands' kernels spin on one another: every band stream needs its own hardware queue, else a waiting sync
        # kernel blocks the launches of the band it waits for (the variable is read when CUDA initialises)
        if nranks > 1 and int(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8")) < 2 * nranks:
            raise RuntimeError(f"LocalBands({nranks}) needs CUDA_DEVICE_MAX_CONNECTIONS >= {2 * nranks} set BEFORE CUDA is "
                               "initialised (tests/conftest.py and bench.py set 32)")
        nbytes = RowTiledRawFormer.comm_bytes(model, H, W, nranks)
        self.regions = [_CommRegion(nbytes, dev) for _ in range(nranks)]
        ptrs = [r.ptr for r in self.regions]
        self.streams = [torch.cuda.Stream(dev) for _ in range(nranks)]
        self.device = dev
        model.packed_weights(dev, _lib.RF_BF16)
        torch.cuda.current_stream(dev).synchronize()
        self.ranks = []
        for r, s in enumerate(self.streams):
            with torch.cuda.stream(s):       # the rehearsal also warms this stream's allocator pool
                self.ranks.append(RowTiledRawFormer(model, H, W, r, nranks, ptrs))
        if graphs:
            for t in self.ranks:
                t.enable_cuda_graphs()

    @torch.no_grad()
    def forward(self, raw):
        cur = torch.cuda.current_stream(self.device)
        self.ranks[0].model.packed_weights(raw.device, _lib.RF_BF16)     # pack once, on the current stream
        outs = []
        if self.ranks[0]._graphs is not None:
            raw = self.ranks[0].model._check_input(raw)
            for t in self.ranks:         # capturing synchronises the device: do it before any band is in flight
                t.capture(raw)
        for t, s in zip(self.ranks, self.streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                o = t(raw)
                o.record_stream(cur)
                outs.append(o)
        for s in self.streams:
            cur.wait_stream(s)
        for t in self.ranks:
            t.status()
        return torch.cat(outs, dim=2)

    __call__ = forward

    def close(self):
        for r in self.regions:
            r.free()
        self.regions = []
