"""N > 1 host logic on CPU: world_size-2 gloo.  The hot path shards by frame (image-parallel, no data-path collective);
what the ranks share is only the bookkeeping bench.py does: per-rank frame assignment, max-over-ranks timing and the
whole-job aggregate.  No GPU needed."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard_frames(n_frames, rank, world):
    """frame i -> rank i mod world (image-parallel)."""
    return list(range(rank, n_frames, world))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_frames(8, rank, world)
    # every frame is owned exactly once
    owned = torch.zeros(8, dtype=torch.int64)
    owned[mine] = 1
    dist.all_reduce(owned)
    # max-over-ranks timing, sum-over-ranks work (bench.py aggregate)
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    frames = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(frames)
    dist.barrier()
    q.put((rank, owned.tolist(), float(ms.item()), float(frames.item())))
    dist.destroy_process_group()


def test_image_parallel_bookkeeping_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, owned, ms, frames in res:
        assert owned == [1] * 8
        assert ms == 15.0          # max over ranks
        assert frames == 8.0       # whole-job aggregate


def test_shard_is_a_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_frames(13, r, world))
        assert seen == list(range(13))
