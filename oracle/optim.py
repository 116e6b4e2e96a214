"""TEST INFRASTRUCTURE (oracle): CPU restatement of the optimiser step around the path.

main.lua:122,133 calls `optim.rmsprop(eval_objective_grad, weights, rmsprop_state)` with
rmsprop_state = {learningRate = opt.lr, alpha = opt.rms_decay}; objective.lua:200 divides the flat gradient by
cls_count before returning it.  `optim` is an un-vendored Torch7 package with no pinned version, so this
restates its published algorithm (optim/rmsprop.lua as of 2015):

    local lr = config.learningRate or 1e-2; local alpha = config.alpha or 0.99
    local epsilon = config.epsilon or 1e-8; local wd = config.weightDecay or 0
    if wd ~= 0 then dfdx:add(wd, x) end
    state.m = state.m or zeros;  state.m:mul(alpha);  state.m:addcmul(1.0 - alpha, dfdx, dfdx)
    state.tmp:sqrt(state.m):add(epsilon)
    x:addcdiv(-lr, dfdx, state.tmp)

Every statement is a TH vector op on FloatTensors (main.lua:51): each elementary operation is rounded to fp32, scalar
arguments are Lua doubles cast to float at the call (so 1.0 - alpha is evaluated in double first).  numpy float32
arithmetic reproduces that exactly.  PARITY UNPINNED: no reference test or golden vector exists for this step."""
import numpy as np

f32 = np.float32


def gradient_div(gradient, cls_count):
    """objective.lua:200: gradient:div(cls_count)."""
    return (gradient.astype(f32) / f32(cls_count)).astype(f32)


def rmsprop_step(x, dfdx, state, learningRate=1e-2, alpha=0.99, epsilon=1e-8, weightDecay=0.0):
    """One optim.rmsprop update.  x, dfdx: float32 arrays (x is updated in place and returned); state: dict with 'm'."""
    x = x.astype(f32, copy=False)
    d = dfdx.astype(f32, copy=True)
    if weightDecay != 0:
        d = (d + f32(weightDecay) * x).astype(f32)                      # dfdx:add(wd, x)
    if "m" not in state:
        state["m"] = np.zeros_like(x, dtype=f32)
    m = state["m"]
    m *= f32(alpha)                                                      # m:mul(alpha)
    m += ((f32(1.0 - alpha) * d).astype(f32) * d).astype(f32)            # m:addcmul(1 - alpha, dfdx, dfdx)
    tmp = (np.sqrt(m).astype(f32) + f32(epsilon)).astype(f32)            # tmp:sqrt(m):add(epsilon)
    x += ((f32(-learningRate) * d).astype(f32) / tmp).astype(f32)        # x:addcdiv(-lr, dfdx, tmp)
    return x
