"""Data-parallel check + timing of the C-ABI gradient all-reduce (frcnn_dp_*), one process per GPU under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
1. the library's in-place bucketed all-reduce gives bit-for-bit what torch.distributed's all_reduce gives on the same
   per-rank gradients (a sum over ranks in NCCL's fixed order), with and without overlap;
2. step time of lossAndGradient (configs[2]: 8 frames 800x450 per GPU) with the all-reduce off / after backward / overlapped:
   the exposed time of the collective.  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frcnn_b200 as F  # noqa: E402
from oracle import anchors as OA, model as OM, objective as OO  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    desc, cfg = OM.VGG_SMALL, OM.CFG_DUPLO
    params = OM.init_params(desc, cfg, seed=0, randomize_aux=True)
    h, w, B = 450, 800, int(os.environ.get("DP_FRAMES", "8"))
    m = F.vgg_small(F.duplo_cfg, device=local)
    m.load_params(params)
    oa = OA.Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])
    dims = m.output_dims(h, w)
    batch = []
    for s in range(B):
        pos, neg, _ = OO.synthetic_examples(oa, dims, w, h, 128, 128, 8, cfg["class_count"], seed=1000 * rank + s)
        pos, neg = F.clean_anchors(pos, dims), F.clean_anchors(neg, dims)
        batch.append(dict(img=OM.synthetic_frame(h, w, seed=100 * rank + s).cuda(), positive=pos, negative=neg,
                          packed=(m.pack_examples(pos), m.pack_examples(neg))))
    local_only = F.create_objective(m, None, defer_div=True, rank=rank)   # no collective; this rank's dropout seeds
    lib_dp = F.create_objective(m, dist, defer_div=True)               # frcnn_dp_allreduce (overlapped: one size group)
    # ---- 1. equality with torch.distributed on IDENTICAL per-rank gradients.  The backward pass itself is not bit
    # reproducible run to run (ROI-pool scatter atomics, TMA reduce-add split-K weight gradients), so the same local
    # gradient buffer is reduced both ways: by torch.distributed on a clone, by the library in place.
    L = F.lib()
    local_only(batch, seed=7)
    g0 = m.gradient.clone()
    ref = g0.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    L.frcnn_dp_set_overlap(m.ctx, 0)
    F.dp_allreduce(m, [0.0] * 7)
    same_plain = bool(torch.equal(m.gradient, ref))
    # the overlapped path sends its buckets from inside the backward pass: a fresh backward, so it is compared with the
    # reduced gradient of the other run to the run-to-run reproducibility of the backward pass itself
    _, g, st = lib_dp(batch, seed=7)
    rel = ((g - ref).norm() / ref.norm()).item()
    local_only(batch, seed=7)
    rel_rerun = ((m.gradient - g0).norm() / g0.norm()).item()
    same_overlap = dict(rel_l2_vs_other_run=rel, rel_l2_of_two_local_runs=rel_rerun, ok=bool(rel <= 4 * rel_rerun + 1e-6))
    cnt = torch.tensor([float(sum(len(b["positive"]) + len(b["negative"]) for b in batch))], device="cuda")
    dist.all_reduce(cnt)
    counts_ok = abs(st["cls_count"] - cnt.item()) < 0.5

    # ---- 2. timing
    def timed(fn, steps=20, warm=5):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(100 + i)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def step_after(i):   # all-reduce strictly after the backward pass
        local_only(batch, seed=i)
        F.dp_allreduce(m, [0.0] * 7)

    ms_none = timed(lambda i: local_only(batch, seed=i))
    L.frcnn_dp_set_overlap(m.ctx, 0)
    ms_after = timed(step_after)
    ms_overlap = timed(lambda i: lib_dp(batch, seed=i))
    nbytes = m.gradient.numel() * 4
    ver = F.ffi.new("int*")
    L.frcnn_dp_info(m.ctx, F.ffi.NULL, F.ffi.NULL, ver, F.ffi.NULL)
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, frames_per_gpu=B, gradient_mb=nbytes / 1e6, nccl_version=ver[0],
                              equal_to_torch_allreduce=dict(overlapped=same_overlap, after_backward=same_plain, counters=counts_ok),
                              ms_per_step=dict(no_collective=ms_none, allreduce_after_backward=ms_after, allreduce_overlapped=ms_overlap),
                              exposed_ms=dict(after_backward=ms_after - ms_none, overlapped=ms_overlap - ms_none),
                              images_per_sec=dict(no_collective=world * B / ms_none * 1e3, overlapped=world * B / ms_overlap * 1e3),
                              busbw_gbs_after=2 * (world - 1) / world * nbytes / max(ms_after - ms_none, 1e-6) / 1e6)), flush=True)
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
