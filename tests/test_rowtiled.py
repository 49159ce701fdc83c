"""Row-tiled single frame (BASELINE config 4).

CPU (``-m "not gpu"``): the band plan, the comm-region / workspace sizing through the C ABI (dry runs, no compute) and
the world_size-2 gloo handshake that hands the IPC handles round.
GPU (``-m gpu``): all bands of a frame as concurrent streams of one B200 (``LocalBands``: the same kernels and the same
flag/mailbox protocol as the multi-GPU run) against (a) the whole-frame forward of the same engine and (b) the CPU
oracle (``oracle/rawformer_torch.py``, pinned to the reference by tests/test_oracle_golden.py)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

import rf_testlib as T


# ---------------------------------------------------------------------------------------------------------
# CPU: plan and sizing
# ---------------------------------------------------------------------------------------------------------
def test_plan_bands_partitions_the_frame():
    from bayer_low_light_image_enhancement_b200 import plan_bands

    for H in (64, 128, 256, 1024, 2848):
        for n in range(1, 9):
            if H // 16 < 4 * n:
                with pytest.raises(ValueError):
                    plan_bands(H, n)
                continue
            bands = plan_bands(H, n)
            assert bands[0][0] == 0 and bands[-1][0] + bands[-1][1] == H
            for (r0, rows), (r1, _) in zip(bands, bands[1:]):
                assert r0 + rows == r1
            assert all(r0 % 16 == 0 and rows % 16 == 0 and rows >= 64 for r0, rows in bands)
            assert max(r for _, r in bands) - min(r for _, r in bands) <= 16
    # 2848 rows = 178 units of 16
    assert [r // 16 for _, r in plan_bands(2848, 2)] == [89, 89]
    assert sorted(r // 16 for _, r in plan_bands(2848, 4)) == [44, 44, 45, 45]
    assert sorted(r // 16 for _, r in plan_bands(2848, 8)) == [22] * 6 + [23] * 2
    with pytest.raises(ValueError):
        plan_bands(2840, 2)
    with pytest.raises(ValueError):
        plan_bands(2848, 9)


def test_band_sizing_through_the_abi():
    """Dry runs of the band plan: the comm region is the same for every rank, the workspace shrinks with the band."""
    from bayer_low_light_image_enhancement_b200 import _lib, plan_bands

    lib = _lib.load()
    H, W = 2848, 4256
    whole = lib.rf_rawformer_workspace_bytes(64, _lib.RF_BF16, 0, 1, H, W)
    for n in (2, 4, 8):
        comm = lib.rf_band_comm_bytes(64, _lib.RF_BF16, 0, H, W, n)
        assert 4096 < comm < 64 << 20
        for r, (row0, rows) in enumerate(plan_bands(H, n)):
            b = _lib.Band()
            b.rank, b.nranks, b.row0, b.rows = r, n, row0, rows
            ws = lib.rf_rawformer_band_workspace_bytes(64, _lib.RF_BF16, 0, H, W, C.byref(b))
            assert 0 < ws < whole * (1.0 / n + 0.2)
            orows, r0 = C.c_int(0), C.c_int(0)
            assert lib.rf_band_out_rows(C.byref(b), C.byref(orows), C.byref(r0)) == 0
            assert r0.value == (8 if r > 0 else 0)
            assert orows.value == rows + r0.value + (8 if r < n - 1 else 0)
    # the multi-level variant has one more sync point (the colour anchor's output sums) -> a slightly larger region
    assert lib.rf_band_comm_bytes(64, _lib.RF_BF16, 1, H, W, 2) > lib.rf_band_comm_bytes(64, _lib.RF_BF16, 0, H, W, 2)
    # unsupported: fp32 engine, bands that do not tile the frame
    assert lib.rf_band_comm_bytes(64, _lib.RF_F32, 0, H, W, 2) == 0
    assert lib.rf_band_comm_bytes(32, _lib.RF_BF16, 0, 256, 256, 5) == 0
    b = _lib.Band()
    b.rank, b.nranks, b.row0, b.rows = 0, 2, 16, 1424          # rank 0 must start at row 0
    assert lib.rf_rawformer_band_workspace_bytes(64, _lib.RF_BF16, 0, H, W, C.byref(b)) == 0
    b.row0, b.rows = 0, 1430                                    # not a multiple of 16
    assert lib.rf_rawformer_band_workspace_bytes(64, _lib.RF_BF16, 0, H, W, C.byref(b)) == 0


def test_row_tiled_rejects_what_it_does_not_implement():
    """No silent fallback: CPU models and the fp32 engine are refused before any device work."""
    import bayer_low_light_image_enhancement_b200 as rf

    with pytest.raises(RuntimeError):                                   # model on the CPU
        rf.RowTiledRawFormer(rf.RawFormer(dim=32, precision="bf16"), 128, 128, 0, 2, [0, 0])
    with pytest.raises(RuntimeError):                                   # ML_RF variant: supported, but not on the CPU either
        rf.RowTiledRawFormer(rf.multilevel.RawFormer(dim=32, precision="bf16"), 128, 128, 0, 2, [0, 0])
    with pytest.raises(ValueError):
        rf.plan_bands(100, 2)
    if torch.cuda.is_available():
        with pytest.raises(NotImplementedError):                        # fp32 parity engine
            rf.RowTiledRawFormer(rf.RawFormer(dim=32, precision="fp32").cuda(), 128, 128, 0, 2, [0, 0])


def _handshake_worker(rank, world, port, q):
    import torch.distributed as dist

    from bayer_low_light_image_enhancement_b200 import plan_bands

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what RowTiledRawFormer.from_process_group exchanges: one 64-byte IPC handle per rank, every rank sees all of them
    mine = bytes([rank + 1]) * 64
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    row0, rows = plan_bands(256, world)[rank]
    cover = torch.zeros(256, dtype=torch.int64)
    cover[row0:row0 + rows] = 1
    dist.all_reduce(cover)
    dist.barrier()
    q.put((rank, [h[0] for h in handles], int(cover.min()), int(cover.max())))
    dist.destroy_process_group()


def test_handle_exchange_world2_gloo():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_handshake_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, firsts, lo, hi in res:
        assert firsts == [1, 2]          # handles arrive in rank order on every rank
        assert (lo, hi) == (1, 1)        # the bands cover every row exactly once


# ---------------------------------------------------------------------------------------------------------
# GPU: bands as concurrent streams of one device
# ---------------------------------------------------------------------------------------------------------
def _psnr(a, b, data_range):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(data_range ** 2 / mse)


CASES = [
    # dim, H, W, bands, input, weight scale
    (32, 128, 128, 2, "rand", 2.0),
    (32, 256, 192, 3, "dark", 2.0),      # uneven bands: 96 + 80 + 80 rows
    (48, 128, 160, 2, "rand", 1.5),
    (64, 192, 128, 3, "rand", 1.0),
    (32, 64, 64, 1, "rand", 2.0),        # one band = the whole frame (no halo, no peers)
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[f"d{c[0]}_{c[1]}x{c[2]}_n{c[3]}" for c in CASES])
def test_row_tiled_matches_whole_frame_and_oracle(case):
    import bayer_low_light_image_enhancement_b200 as rf
    from oracle import rawformer_torch as oracle

    dim, H, W, n, kind, scale = case
    dev = torch.device("cuda", 0)
    m = rf.RawFormer(dim=dim, precision="bf16")
    sd = T.make_state_dict(m, seed=4321 + dim, scale=scale)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    x = torch.from_numpy(T.gen_input(kind, (1, 1, H, W), 3)).to(dev)
    with torch.no_grad():
        whole = m(x).float().cpu().numpy()
    bands = rf.LocalBands(m, H, W, n)
    try:
        for it in range(2):              # twice: the mailboxes and the epoch counters are reused frame after frame
            tiled = bands(x).float().cpu().numpy()
            assert tiled.shape == whole.shape and np.isfinite(tiled).all()
            rng = max(float(whole.max() - whole.min()), 1e-6)
            # same engine, same arithmetic per pixel; only the order of the fp32 partial sums of the three per-image
            # reductions differs, which moves a few bf16 roundings: far tighter than the bf16-vs-reference bar
            p = _psnr(tiled, whole, rng)
            assert p >= 54.0, f"frame {it}: PSNR(row-tiled, whole-frame) = {p:.1f} dB"
    finally:
        bands.close()
    with torch.no_grad():
        ref = oracle.rawformer_forward({k: v.float() for k, v in sd.items()}, x.cpu(), "flca").numpy()
    rng = max(float(ref.max() - ref.min()), 1e-6)
    p_t, p_w = _psnr(tiled, ref, rng), _psnr(whole, ref, rng)
    assert p_t >= 35.0, f"PSNR(row-tiled, oracle) = {p_t:.1f} dB"
    assert abs(p_t - p_w) <= 0.5, f"row-tiled {p_t:.2f} dB vs whole-frame {p_w:.2f} dB against the oracle"


ML_CASES = [
    # dim, H, W, bands, input, weight scale
    (32, 128, 128, 2, "rand", 1.5),
    (32, 256, 192, 3, "dark", 1.5),      # uneven bands, a band with two halos
    (64, 192, 128, 2, "rand", 1.0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ML_CASES, ids=[f"ml_d{c[0]}_{c[1]}x{c[2]}_n{c[3]}" for c in ML_CASES])
def test_row_tiled_multilevel_matches_whole_frame_and_oracle(case):
    """ML_RF.py::RawFormer row-tiled (config 4 x config 5): FLCA_Pyramid gates from the whole-frame guidance sums (replicated),
    channel sums over the band's interior, colour anchor (ML_RF.py:284-285) with the output-channel sums all-reduced over the
    bands, LL nudge at the band's frame rows."""
    import bayer_low_light_image_enhancement_b200 as rf
    from oracle import rawformer_torch as oracle

    dim, H, W, n, kind, scale = case
    dev = torch.device("cuda", 0)
    m = rf.multilevel.RawFormer(dim=dim, precision="bf16")
    sd = T.make_state_dict(m, seed=977 + dim, scale=scale)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    x = torch.from_numpy(T.gen_input(kind, (1, 1, H, W), 5)).to(dev)
    with torch.no_grad():
        whole = m(x).float().cpu().numpy()
    bands = rf.LocalBands(m, H, W, n)
    try:
        for it in range(2):
            tiled = bands(x).float().cpu().numpy()
            assert tiled.shape == whole.shape and np.isfinite(tiled).all()
            rng = max(float(whole.max() - whole.min()), 1e-6)
            p = _psnr(tiled, whole, rng)
            assert p >= 50.0, f"frame {it}: PSNR(row-tiled, whole-frame) = {p:.1f} dB"
    finally:
        bands.close()
    with torch.no_grad():
        ref = oracle.rawformer_forward({k: v.float() for k, v in sd.items()}, x.cpu(), "ml").numpy()
    rng = max(float(ref.max() - ref.min()), 1e-6)
    p_t, p_w = _psnr(tiled, ref, rng), _psnr(whole, ref, rng)
    assert p_t >= 35.0, f"PSNR(row-tiled, oracle) = {p_t:.1f} dB"
    assert abs(p_t - p_w) <= 0.5, f"row-tiled {p_t:.2f} dB vs whole-frame {p_w:.2f} dB against the oracle"


@pytest.mark.gpu
def test_row_tiled_graph_replay():
    """Every band's forward as one CUDA graph (the frame counter is advanced on the device): same result as eager
    launches, frame after frame, also when the input buffer's content changes.  Not bit-equal: the fp32 atomics of the
    Gram / channel-sum kernels make two runs of the SAME path differ at ~58 dB with these stress weights."""
    import bayer_low_light_image_enhancement_b200 as rf

    dev = torch.device("cuda", 0)
    m = rf.RawFormer(dim=32, precision="bf16")
    m.load_state_dict(T.make_state_dict(m, seed=5, scale=2.0), strict=True)
    m = m.to(dev).eval()
    H, W, n = 192, 128, 3
    x = torch.from_numpy(T.gen_input("rand", (1, 1, H, W), 21)).to(dev)
    x2 = torch.from_numpy(T.gen_input("dark", (1, 1, H, W), 22)).to(dev)
    eager, graphs = rf.LocalBands(m, H, W, n), rf.LocalBands(m, H, W, n, graphs=True)
    try:
        buf = x.clone()
        for frame in (x, x2, x):
            buf.copy_(frame)
            with torch.no_grad():
                whole = m(buf).float().cpu().numpy()
            a = eager(buf).float().cpu().numpy()
            b = graphs(buf).float().cpu().numpy()
            rng = float(whole.max() - whole.min())
            assert _psnr(b, whole, rng) >= 54.0 and _psnr(b, a, rng) >= 54.0, (
                f"graph/whole {_psnr(b, whole, rng):.1f} dB, graph/eager {_psnr(b, a, rng):.1f} dB, "
                f"eager/whole {_psnr(a, whole, rng):.1f} dB")
    finally:
        eager.close()
        graphs.close()


@pytest.mark.gpu
def test_row_tiled_band_seams():
    """The rows next to a band boundary are where a wrong halo shows: compare them separately."""
    import bayer_low_light_image_enhancement_b200 as rf

    dev = torch.device("cuda", 0)
    m = rf.RawFormer(dim=32, precision="bf16")
    m.load_state_dict(T.make_state_dict(m, seed=99, scale=2.0), strict=True)
    m = m.to(dev).eval()
    H, W, n = 256, 128, 4
    x = torch.from_numpy(T.gen_input("rand", (1, 1, H, W), 11)).to(dev)
    with torch.no_grad():
        whole = m(x)
    bands = rf.LocalBands(m, H, W, n)
    try:
        tiled = bands(x)
    finally:
        bands.close()
    rng = float(whole.max() - whole.min())
    err = (tiled - whole).abs().amax(dim=(0, 1, 3)).cpu().numpy() / rng       # per output row
    seams = [r0 for r0, _ in rf.plan_bands(H, n)][1:]
    near = np.zeros(H, bool)
    for s in seams:
        near[s - 16:s + 16] = True
    assert err[near].max() <= 4.0 * max(err[~near].max(), 2e-3), (
        f"seam rows max rel err {err[near].max():.4f} vs interior {err[~near].max():.4f}")
