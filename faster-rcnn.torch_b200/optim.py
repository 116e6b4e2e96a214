"""Host mirror of the optimiser call of main.lua:122,133: rmsprop(opfunc, model, config, state).

One fused CUDA pass (frcnn_rmsprop_step) over the model's flat `weights` / `gradient` buffers and the state buffer `m`
replaces gradient:div(cls_count) (objective.lua:200) and the four TH vector passes of optim.rmsprop; the tensor-core
weight packs are refreshed afterwards (frcnn_pack_weights), as SURVEY 8b requires after every optimiser step."""
import torch

from ._lib import check, ffi, lib


def rmsprop_step(model, state, learningRate=1e-2, alpha=0.99, epsilon=1e-8, weightDecay=0.0, grad_div=1.0, repack=True):
    """Applies one optim.rmsprop update to model.weights from model.gradient.  state: dict holding 'm' (created zeroed on
    first use, like rmsprop_state.m).  grad_div: divide the gradient first (the deferred gradient:div(cls_count))."""
    if "m" not in state:
        state["m"] = torch.zeros_like(model.weights)
    w, g, m = model.weights, model.gradient, state["m"]
    # The BatchNorm running statistics live in the flat buffer here (so that one snapshot carries them), but they are not
    # parameters in Torch (nn.Module:parameters() / getParameters leave them out): their gradient slots stay zero, and the
    # only term that could move them is weightDecay * x -- keep them out of the update.
    stats = [n for n in model.param_names if n.endswith((".bn_mean", ".bn_var"))] if weightDecay != 0.0 else []
    saved = [model.params[n].clone() for n in stats]
    check(model.ctx, lib().frcnn_rmsprop_step(model.ctx, ffi.cast("float*", w.data_ptr()), ffi.cast("float*", g.data_ptr()),
                                              ffi.cast("float*", m.data_ptr()), w.numel(), float(grad_div), float(learningRate),
                                              float(alpha), float(epsilon), float(weightDecay)))
    for n, t in zip(stats, saved):
        model.params[n].copy_(t)
    if repack:
        model.pack_weights()
    return w


def sync_running_stats(model, dist):
    """Data-parallel replicas see different frames, so their BatchNorm running statistics drift apart (the reference is
    single-GPU and has nothing to say here): average them across ranks, e.g. once per snapshot interval."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    names = [n for n in model.param_names if n.endswith((".bn_mean", ".bn_var"))]
    if not names:
        return
    buf = torch.cat([model.params[n].reshape(-1) for n in names])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    buf /= dist.get_world_size()
    off = 0
    for n in names:
        k = model.params[n].numel()
        model.params[n].copy_(buf[off:off + k].reshape(model.params[n].shape))
        off += k


def rmsprop(opfunc, model, config=None, state=None):
    """optim.rmsprop(opfunc, x, config, state): opfunc() -> (loss, gradient[, stats]) evaluates lossAndGradient into
    model.gradient (create_objective); returns (model.weights, [loss]).  With opfunc built by
    create_objective(..., defer_div=True) the division by cls_count is fused into the update."""
    config = config if config is not None else {}
    state = state if state is not None else config
    out = opfunc()
    loss, stats = out[0], (out[2] if len(out) > 2 else {})
    grad_div = float(stats.get("deferred_div", 1.0)) if isinstance(stats, dict) else 1.0
    rmsprop_step(model, state, config.get("learningRate", 1e-2), config.get("alpha", 0.99), config.get("epsilon", 1e-8),
                 config.get("weightDecay", 0.0), grad_div=grad_div)
    return model.weights, [loss]
