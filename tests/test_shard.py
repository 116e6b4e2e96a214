"""CPU tests of the N > 1 host logic: how frames / class segments shard over ranks and how the timing scalars are
reduced, run as a real world_size-2 gloo job (no GPU)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_partitions(F):
    for n, w in ((8, 8), (8, 3), (1, 2), (21, 8), (5, 4)):
        spans = [F.shard_frames(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    got = [len(F.shard_segments(21, 8, r)) for r in range(8)]
    assert got == [3, 3, 3, 3, 3, 2, 2, 2]
    assert sorted(s for r in range(8) for s in F.shard_segments(21, 8, r)) == list(range(21))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import frcnn_b200 as F
    from oracle import boxes as OB, nms as ON
    # the NMS sweep sharded by class segment: every rank runs its own segments (oracle on CPU here; the CUDA
    # library on the GPU box), no data-path collective; only the unit counts / timings are reduced
    n = 3000
    perm, seg = OB.class_segments(n, 21, seed=4)
    boxes = OB.sweep_boxes(n, seed=4)[perm]
    mine = F.shard_segments(21, world, rank)
    picks = {s: ON.nms(boxes[seg[s]:seg[s + 1]], 0.25) for s in mine}
    units = sum(int(seg[s + 1] - seg[s]) for s in mine)
    ms, total = F.reduce_timing(10.0 + rank, units, world, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, {s: p.tolist() for s, p in picks.items()})
    dist.barrier()
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        torch.save(dict(ms=ms, total=total, merged=merged), out)
    dist.destroy_process_group()


def test_two_rank_gloo_job(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ms"] == 11.0 and res["total"] == 3000.0  # max over ranks, sum over ranks
    from oracle import boxes as OB, nms as ON
    perm, seg = OB.class_segments(3000, 21, seed=4)
    boxes = OB.sweep_boxes(3000, seed=4)[perm]
    assert sorted(res["merged"]) == list(range(21))
    for s in range(21):
        assert res["merged"][s] == ON.nms(boxes[seg[s]:seg[s + 1]], 0.25).tolist()


def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import frcnn_b200 as F
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    grad, c = F.allreduce_gradient(g, [1.0 + rank, 2.0, 3.0, 4.0, 100 + rank, 50, 1], dist)
    if rank == 0:
        torch.save(dict(grad=grad, c=c), out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks(tmp_path):
    """The one collective of the path: flat gradient + loss / example counters summed in one all-reduce."""
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, 31500 + os.getpid() % 2000, out), nprocs=2, join=True)
    res = torch.load(out)
    assert torch.equal(res["grad"], torch.arange(10, dtype=torch.float32) * 3)
    assert res["c"].tolist() == [3.0, 4.0, 6.0, 8.0, 201.0, 100.0, 2.0]
    import frcnn_b200 as F
    g = torch.ones(4)
    grad, c = F.allreduce_gradient(g, [1, 2, 3, 4, 5, 6, 7], None)   # single process: identity
    assert torch.equal(grad, torch.ones(4)) and c.tolist() == [1, 2, 3, 4, 5, 6, 7]
