"""Host mirror of objective.lua: create_objective(model, ...) -> lossAndGradient.

The per-anchor Lua loops, the per-ROI pooling calls and every criterion of objective.lua:65-198 run inside
frcnn_train_image (one call per frame); what stays here is what objective.lua does around them: zero the flat
gradient (objective.lua:49), loop over the frames of the batch (:65), sum the statistics, and divide by the number of
anchor examples (:200).  With torch.distributed initialised the frames of a batch are sharded over the ranks and the
flat gradient plus the example / loss counters are summed with ONE all-reduce (NCCL over NVLink on the GPU box)
before the division -- the only collective of the path."""
import torch


def clean_anchors(examples, dims):
    """cleanAnchors (objective.lua:32-43): drop examples whose index lies outside the actual feature map."""
    return [e for e in examples if e[0].index[1] <= dims[e[0].layer - 1][1] and e[0].index[2] <= dims[e[0].layer - 1][2]]


def allreduce_gradient(gradient, counters, dist=None):
    """Sums the flat gradient and the counters [cls_loss, reg_loss, creg_loss, ccls_loss, cls_count, reg_count,
    ccls_count] over the ranks with one collective: the counters ride in the tail of the same buffer."""
    c = torch.as_tensor(counters, dtype=gradient.dtype, device=gradient.device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return gradient, c
    buf = torch.cat([gradient, c])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    gradient.copy_(buf[:gradient.numel()])
    return gradient, buf[gradient.numel():].clone()


def dp_init(model, dist):
    """Forms the C library's NCCL communicator (frcnn_dp_init_rank) over the ranks of an initialised torch.distributed
    group -- which is only used to ship rank 0's 128-byte NCCL id.  Returns the (device) counter buffer the all-reduce
    carries next to the gradient."""
    from ._lib import check, ffi, lib
    L = lib()
    idbuf = ffi.new("char[128]")
    payload = [None]
    if dist.get_rank() == 0:
        check(None, L.frcnn_dp_unique_id(idbuf))
        payload = [bytes(ffi.buffer(idbuf, 128))]
    dist.broadcast_object_list(payload, src=0, device=model.device)
    ffi.memmove(idbuf, payload[0], 128)
    check(model.ctx, L.frcnn_dp_init_rank(model.ctx, idbuf, dist.get_rank(), dist.get_world_size()))
    model._dp_counters = torch.zeros(8, dtype=torch.float32, device=model.device)
    return model._dp_counters


def dp_allreduce(model, counters):
    """frcnn_dp_allreduce on this rank's context: the gradient buckets pnet:backward has not sent yet + the counters, in
    place, on the library's side stream; returns the summed counters (host list)."""
    from ._lib import check, ffi, lib
    buf = model._dp_counters
    buf[:len(counters)].copy_(torch.tensor(counters, dtype=torch.float32))
    ctxs = ffi.new("frcnn_ctx*[]", [model.ctx])
    ptrs = ffi.new("float*[]", [ffi.cast("float*", buf.data_ptr())])
    check(model.ctx, lib().frcnn_dp_allreduce(ctxs, 1, ptrs, len(counters)))
    return buf[:len(counters)].tolist()


def create_objective(model, dist=None, defer_div=False, batched=True, rank=None):
    """Returns lossAndGradient(batch, seed) -> (loss, gradient, stats); batch = list of dicts {img [3][H][W] tensor,
    positive [(anchor, roi)], negative [(anchor,)]} as BatchIterator:nextTraining yields them (this rank's share).
    batched: frames of equal size are processed by one frcnn_train_batch call (False: frame by frame, frcnn_train_image).
    defer_div: leave gradient:div(cls_count) (objective.lua:200) to the fused optimiser pass (optim.rmsprop_step's
    grad_div = stats['deferred_div']).  rank: the rank the dropout seeds are derived from (default: dist's rank, 0 without
    a group) -- a local objective that must draw the masks of a given rank passes it here."""

    # With NCCL available the collective runs inside the C library (frcnn_dp_*: in place, bucketed, overlapped with
    # pnet:backward); the torch.distributed path below remains for host-only groups (gloo tests).
    use_lib_dp = (dist is not None and dist.is_initialized() and dist.get_world_size() > 1 and not model.host_only
                  and dist.get_backend() == "nccl")
    if use_lib_dp and getattr(model, "_dp_counters", None) is None:
        dp_init(model, dist)
    from ._lib import lib as _lib

    counter = dict(step=0)
    if rank is None:
        rank = dist.get_rank() if (dist is not None and dist.is_initialized()) else 0

    def lossAndGradient(batch, seed=None):
        # the reference draws fresh SpatialDropout / Dropout masks from the global generator on every forward: without an
        # explicit seed every call gets a new one, and every rank / frame its own (seed mixes step, rank and frame index)
        counter["step"] += 1
        if seed is None:
            seed = counter["step"]
        frame0 = rank * len(batch)
        model.zero_grad()                                   # gradient:zero()
        model.pnet.training()
        model.cnet.training()
        sums = dict(cls=0.0, reg=0.0, creg=0.0, ccls=0.0)
        cls_count = reg_count = ccls_count = 0
        # frames of one size go through frcnn_train_batch together (pnet forward / backward once over all of them);
        # frames of different sizes one by one, as the reference's loop does
        groups = {}
        for i, x in enumerate(batch):
            groups.setdefault(tuple(x["img"].shape), []).append(i)
        if use_lib_dp:
            # one size group = one frcnn_train_batch call = the step's only accumulation into the gradient: its buckets may
            # leave as soon as the backward pass has finished them
            _lib().frcnn_dp_set_overlap(model.ctx, 1 if (len(groups) == 1 and batched) else 0)
        elif getattr(model, "_dp_counters", None) is not None:
            _lib().frcnn_dp_set_overlap(model.ctx, 0)       # a local objective on a context that owns a communicator
        for shape, idx in groups.items():
            dims = model.output_dims(shape[1], shape[2])
            ps = [clean_anchors(batch[i]["positive"], dims) for i in idx]
            ns = [clean_anchors(batch[i]["negative"], dims) for i in idx]
            packed = None
            # example records marshalled once by the caller (BatchIterator) -- only usable when cleanAnchors
            # (objective.lua:32-43) dropped nothing: the records are index-aligned with the uncleaned lists
            if all("packed" in batch[i] for i in idx) and all(len(p) == len(batch[i]["positive"]) and len(q) == len(batch[i]["negative"])
                                                              for p, q, i in zip(ps, ns, idx)):
                packed = [batch[i]["packed"] for i in idx]
            if len(idx) == 1 and not batched:
                all_losses = [model.train_image(batch[idx[0]]["img"], ps[0], ns[0], seed=seed * 1000003 + frame0 + idx[0])]
            else:
                all_losses = model.train_batch([batch[i]["img"] for i in idx], ps, ns,
                                               seeds=[seed * 1000003 + frame0 + i for i in idx], packed=packed)
            for losses, p, n in zip(all_losses, ps, ns):
                for k in sums:
                    sums[k] += losses[k]
                reg_count += len(p)
                cls_count += len(p) + len(n)
                ccls_count += 1
        counters = [sums["cls"], sums["reg"], sums["creg"], sums["ccls"], cls_count, reg_count, ccls_count]
        if use_lib_dp:
            gradient, c = model.gradient, dp_allreduce(model, counters)
        else:
            gradient, c = allreduce_gradient(model.gradient, counters, dist)
            c = c.tolist()
        if not defer_div:
            gradient.div_(max(c[4], 1.0))                    # gradient:div(cls_count)
        stats = dict(pcls=c[0] / max(c[4], 1.0), preg=c[1] / max(c[5], 1.0), dcls=c[3] / max(c[6], 1.0), dreg=c[2] / max(c[5], 1.0),
                     cls_count=int(c[4]), reg_count=int(c[5]), deferred_div=max(c[4], 1.0) if defer_div else 1.0)
        return stats["pcls"] + stats["preg"], gradient, stats

    return lossAndGradient
