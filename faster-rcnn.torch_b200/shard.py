"""How the path shards over the GPUs of one box: frames -- and, for the NMS sweep, class segments --
are independent units, so every rank gets a contiguous / round-robin share and there is NO data-path collective.
torch.distributed is only used for the barrier around the timed region and for reducing the timing scalars
(max over ranks) and the unit counts (sum over ranks)."""


def shard_frames(n_frames, world, rank):
    """Contiguous share of a global batch: the first (n_frames % world) ranks get one extra frame."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_segments(n_seg, world, rank):
    """Class segments of the per-class NMS (Detector.lua:125-136) dealt round-robin: 21 classes on 8 GPUs ->
    3,3,3,3,3,2,2,2."""
    return [s for s in range(n_seg) if s % world == rank]


def reduce_timing(ms_local, units_local, world, dist=None, device=None):
    """(max over ranks of the device time, sum over ranks of the processed units)."""
    if world == 1:
        return float(ms_local), float(units_local)
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
