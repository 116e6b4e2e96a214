-- frcnn_b200.lua -- LuaJIT-FFI glue between the reference's Lua surface and libfrcnn_b200.so.
--
-- NOT RUN IN THIS REPOSITORY'S CI: the build image has no Lua / LuaJIT / Torch7 (SURVEY.md section 0.4).  The file is
-- the reference-side binding a maintainer adds (INTEGRATION.md); it is kept thin and mechanical: every numeric
-- step happens inside the C library, whose ABI (include/frcnn_b200.h) is exercised by the Python cffi tests with
-- the very same cdef text.
--
-- Usage (main.lua stays unchanged except for three lines, see INTEGRATION.md):
--   local b200 = require 'frcnn_b200'
--   local model = load_model(cfg, opt.model, opt.restore, true)     -- main.lua:80-101, unchanged
--   b200.accelerate(model)                                          -- binds the flat weights, builds the plan
--   local d = Detector(model)                                       -- lua/Detector.lua of this directory
local ffi = require 'ffi'
local M = {}

-- The header is cdef-clean: strip the preprocessor lines and the extern "C" braces, pass the rest verbatim.
local function read_header(path)
  local f = assert(io.open(path, 'r'))
  local out = {}
  for line in f:lines() do
    local s = line:match('^%s*(.-)%s*$')
    if s:sub(1, 1) ~= '#' and s ~= 'extern "C" {' and s ~= '}' then out[#out + 1] = line end
  end
  f:close()
  return table.concat(out, '\n')
end

local root = os.getenv('FRCNN_B200_ROOT') or '.'
ffi.cdef(read_header(root .. '/include/frcnn_b200.h'))
local C = ffi.load(root .. '/faster-rcnn.torch_b200/libfrcnn_b200.so')
M.C = C

local function check(ctx, rc)
  if rc ~= 0 then
    error(string.format('frcnn_b200 error %d: %s', rc, ffi.string(C.frcnn_last_error(ctx))), 3)
  end
end
M.check = check

-- One context per device (cutorch.setDevice(opt.gpuid + 1), main.lua:52).  stream nil = legacy default stream, which
-- keeps the ordering with surrounding cutorch work.
function M.create(device)
  local p = ffi.new('frcnn_ctx*[1]')
  check(nil, C.frcnn_create(p, device or (cutorch.getDevice() - 1), nil))
  return ffi.gc(p[0], C.frcnn_destroy)
end

-- Walks the nngraph exactly like Localizer.lua:19-24 does: output node -> children[1] -> ... collecting the
-- nn.Sequential of every conv block (model_utilities.lua:41-49), in forward order.
local function trunk_blocks(pnet, n_heads)
  local blocks, node = {}, pnet.outnode.children[n_heads + 1]
  while node and node.data.module do
    if torch.typename(node.data.module) == 'nn.Sequential' then table.insert(blocks, 1, node.data.module) end
    node = node.children[1]
  end
  return blocks
end

local function dev_ptr(t) return ffi.cast('const float*', t:data()) end

-- Builds the plan from the model's own description tables (models/vgg_*.lua) and binds device pointers of the
-- learnable tensors -- views into the flat CudaTensor created by combine_and_flatten_parameters
-- (utilities.lua:136-147), so optimiser updates (main.lua:133) are seen after the next M.pack(model).
function M.accelerate(model, anchor_nets, class_layers)
  local cfg, layers = model.cfg, model.layers
  anchor_nets = anchor_nets or model.anchor_nets
  class_layers = class_layers or model.class_layers
  assert(anchor_nets and class_layers, 'pass the anchor_nets / class_layers tables of models/vgg_*.lua')
  local ctx = M.create()
  local nb, nh, nf = #layers, #anchor_nets, #class_layers
  local blocks = ffi.new('frcnn_block_desc[?]', nb)
  for i, l in ipairs(layers) do
    local b = blocks[i - 1]
    b.filters, b.kW, b.kH, b.padW, b.padH, b.conv_steps, b.dropout = l.filters, l.kW, l.kH, l.padW, l.padH, l.conv_steps, l.dropout or 0
  end
  local heads = ffi.new('frcnn_head_desc[?]', nh)
  for i, a in ipairs(anchor_nets) do heads[i - 1].kW, heads[i - 1].n, heads[i - 1].input = a.kW, a.n, a.input end
  local fcs = ffi.new('frcnn_fc_desc[?]', nf)
  for i, l in ipairs(class_layers) do
    fcs[i - 1].n, fcs[i - 1].dropout, fcs[i - 1].batch_norm = l.n, l.dropout or 0, l.batch_norm and 1 or 0
  end
  local scales = ffi.new('double[?]', #cfg.scales, cfg.scales)
  check(ctx, C.frcnn_model_plan(ctx, blocks, nb, heads, nh, fcs, nf, cfg.class_count, cfg.roi_pooling.kh, cfg.roi_pooling.kw,
                                scales, #cfg.scales, -1))
  -- parameter pointers in the library's bind order (frcnn_param_info): trunk convs, heads, cnet
  local ptrs = {}
  for _, seq in ipairs(trunk_blocks(model.pnet, nh)) do
    local convs, prelus = seq:findModules('nn.SpatialConvolution'), seq:findModules('nn.PReLU')
    for i = 1, #convs do
      ptrs[#ptrs + 1] = dev_ptr(convs[i].weight); ptrs[#ptrs + 1] = dev_ptr(convs[i].bias); ptrs[#ptrs + 1] = dev_ptr(prelus[i].weight)
    end
  end
  for i = 1, nh do
    local seq = model.pnet.outnode.children[i].data.module   -- AnchorNetwork Sequential (model_utilities.lua:29-35)
    local c1, pr, c2 = seq.modules[1], seq.modules[2], seq.modules[3]
    for _, t in ipairs{c1.weight, c1.bias, pr.weight, c2.weight, c2.bias} do ptrs[#ptrs + 1] = dev_ptr(t) end
  end
  local lin = model.cnet:findModules('nn.Linear')
  local bns = model.cnet:findModules('nn.BatchNormalization')
  local prs = model.cnet:findModules('nn.PReLU')
  local bi = 0
  for i, l in ipairs(class_layers) do
    ptrs[#ptrs + 1] = dev_ptr(lin[i].weight); ptrs[#ptrs + 1] = dev_ptr(lin[i].bias)
    if l.batch_norm then
      bi = bi + 1
      local bn = bns[bi]
      for _, t in ipairs{bn.weight, bn.bias, bn.running_mean, bn.running_var} do ptrs[#ptrs + 1] = dev_ptr(t) end
    end
    ptrs[#ptrs + 1] = dev_ptr(prs[i].weight)
  end
  -- the two output branches: Linear(n, 4) and Linear(n, classes) (model_utilities.lua:96-104)
  local reg, cls = lin[nf + 1], lin[nf + 2]
  if reg.weight:size(1) ~= 4 then reg, cls = cls, reg end
  for _, t in ipairs{reg.weight, reg.bias, cls.weight, cls.bias} do ptrs[#ptrs + 1] = dev_ptr(t) end
  assert(#ptrs == C.frcnn_param_count(ctx), 'parameter walk does not match the plan')
  local arr = ffi.new('const float*[?]', #ptrs, ptrs)
  check(ctx, C.frcnn_bind_params(ctx, arr, #ptrs))
  model.b200 = { ctx = ctx, n_heads = nh }
  M.pack(model)
  return model
end

-- Call after every optimiser step / weights:copy (main.lua:97,133): re-packs fp32 -> bf16 tensor-core layouts.
function M.pack(model)
  cutorch.synchronize()
  check(model.b200.ctx, C.frcnn_pack_weights(model.b200.ctx))
end

-- pnet:forward(img) drop-in (Detector.lua:33): returns the 5-entry output table of CudaTensors.
function M.pnet_forward(model, img)
  local ctx, nh = model.b200.ctx, model.b200.n_heads
  local h, w = img:size(2), img:size(3)
  local dims = ffi.new('int[?]', 3 * (nh + 1))
  check(ctx, C.frcnn_pnet_output_dims(ctx, h, w, dims))
  local outs, ptrs = {}, ffi.new('float*[?]', nh + 1)
  for i = 0, nh do
    outs[i + 1] = torch.CudaTensor(dims[3 * i], dims[3 * i + 1], dims[3 * i + 2])
    ptrs[i] = outs[i + 1]:data()
  end
  check(ctx, C.frcnn_pnet_forward(ctx, img:data(), 1, h, w, ptrs))
  return outs
end

-- ---------------------------------------------------------------------------------------------- training
-- Binds the flat gradient views (same walk as accelerate, but over gradWeight / gradBias).  Call once after
-- combine_and_flatten_parameters (main.lua:92).
function M.bind_grads(model, grad_ptrs)
  local ctx = model.b200.ctx
  local arr = ffi.new('float*[?]', #grad_ptrs, grad_ptrs)
  check(ctx, C.frcnn_bind_grads(ctx, arr, #grad_ptrs))
end

-- The body of the per-image loop of lossAndGradient (objective.lua:65-198) as ONE call: positives = { {anchor, roi}, ... },
-- negatives = { {anchor}, ... } exactly as BatchIterator:nextTraining yields them (after cleanAnchors).
-- Returns cls_loss, reg_loss, creg_loss, ccls_loss of the frame; gradients accumulate in the flat gradient tensor.
function M.train_image(model, img, positives, negatives, seed)
  local ctx = model.b200.ctx
  local function fill(e, anchor, roi)
    e.anchor[0], e.anchor[1], e.anchor[2], e.anchor[3] = anchor.minX, anchor.minY, anchor.maxX, anchor.maxY
    e.layer, e.aspect, e.y, e.x = anchor.layer, anchor.aspect, anchor.index[2], anchor.index[3]
    if roi then
      local r = roi.rect
      e.roi[0], e.roi[1], e.roi[2], e.roi[3] = r.minX, r.minY, r.maxX, r.maxY
      local t = Anchors.inputToAnchor(anchor, r)          -- FloatTensor(4), objective.lua:110
      for k = 0, 3 do e.reg_target[k] = t[k + 1] end
      e.class_index = roi.class_index
    end
  end
  local pos = ffi.new('frcnn_example[?]', math.max(#positives, 1))
  local neg = ffi.new('frcnn_example[?]', math.max(#negatives, 1))
  for i, x in ipairs(positives) do fill(pos[i - 1], x[1], x[2]) end
  for i, x in ipairs(negatives) do fill(neg[i - 1], x[1], nil) end
  local losses = ffi.new('float[4]')
  check(ctx, C.frcnn_train_image(ctx, img:data(), img:size(2), img:size(3), pos, #positives, neg, #negatives, nil, nil,
                                 seed or 0, losses))
  return losses[0], losses[1], losses[2], losses[3]
end

-- Anchors:findPositive (Anchors.lua:147-195) as one launch.  `anchors` is the reference's own Anchors object (its :get
-- builds the returned rects); returns { {anchor_rect, roi}, ... } in the reference's order.
function M.find_positive(model, anchors, roi_list, clip_rect, pos_threshold, neg_threshold, include_best)
  local ctx, n = model.b200.ctx, #roi_list
  if n == 0 then return {} end
  local rois = ffi.new('double[?]', 4 * n)
  for i, roi in ipairs(roi_list) do
    local r = roi.rect
    rois[4 * i - 4], rois[4 * i - 3], rois[4 * i - 2], rois[4 * i - 1] = r.minX, r.minY, r.maxX, r.maxY
  end
  local clip = nil
  if clip_rect then clip = ffi.new('double[4]', clip_rect.minX, clip_rect.minY, clip_rect.maxX, clip_rect.maxY) end
  local cap = 65536
  local out, out_roi, cnt = ffi.new('frcnn_anchor_ref[?]', cap), ffi.new('int[?]', cap), ffi.new('int[1]')
  check(ctx, C.frcnn_find_positive(ctx, rois, n, clip, pos_threshold, neg_threshold, include_best and 1 or 0, out, out_roi, cap, cnt))
  local matches = {}
  for i = 0, cnt[0] - 1 do
    matches[i + 1] = { anchors:get(out[i].layer, out[i].aspect, out[i].y, out[i].x), roi_list[out_roi[i] + 1] }
  end
  return matches
end

-- Anchors:sampleNegative (Anchors.lua:197-235).  The generator stays in Lua: three torch.random() values per trial are
-- drawn here and handed over.  The reference consumes exactly three values per trial it runs, so the generator is put
-- back and advanced by 3 * used afterwards: the Lua random stream stays identical to the reference's.  If the drawn
-- stream runs out before the loop's stopping rule fires (more than 500 rejections spread between the accepts) the loop
-- is continued with the remaining count and the carried run of rejections.
function M.sample_negative(model, anchors, image_rect, roi_list, neg_threshold, count)
  local ctx, n = model.b200.ctx, #roi_list
  local rois = ffi.new('double[?]', math.max(4 * n, 1))
  for i, roi in ipairs(roi_list) do
    local r = roi.rect
    rois[4 * i - 4], rois[4 * i - 3], rois[4 * i - 2], rois[4 * i - 1] = r.minX, r.minY, r.maxX, r.maxY
  end
  local img = ffi.new('double[4]', image_rect.minX, image_rect.minY, image_rect.maxX, image_rect.maxY)
  local neg, out = {}, ffi.new('frcnn_anchor_ref[?]', math.max(count, 1))
  local cnt, used, fin, retry = ffi.new('int[1]'), ffi.new('int[1]'), ffi.new('int[1]'), ffi.new('int[1]')
  local need, carried = count, 0
  repeat
    local trials = need + 500
    local rnd = ffi.new('uint32_t[?]', 3 * trials)
    local rng_state = torch.getRNGState()
    for i = 0, 3 * trials - 1 do rnd[i] = torch.random() end
    check(ctx, C.frcnn_sample_negative(ctx, img, rois, n, neg_threshold, need, rnd, trials, carried, out, need, cnt, used, fin, retry))
    torch.setRNGState(rng_state)
    for i = 1, 3 * used[0] do torch.random() end
    for i = 0, cnt[0] - 1 do neg[#neg + 1] = { anchors:get(out[i].layer, out[i].aspect, out[i].y, out[i].x) } end
    need, carried = need - cnt[0], retry[0]
  until fin[0] == 1 or need <= 0
  return neg
end

-- gradient:div(cls_count) (objective.lua:200) + optim.rmsprop (main.lua:122,133) as one fused pass over the flat
-- CudaTensors; state.m is created zeroed on first use like optim.rmsprop does.
function M.rmsprop_step(model, weights, gradient, state, cls_count)
  state.m = state.m or weights.new(weights:size()):zero()
  check(model.b200.ctx, C.frcnn_rmsprop_step(model.b200.ctx, weights:data(), gradient:data(), state.m:data(), weights:nElement(),
                                             cls_count or 1, state.learningRate or 1e-2, state.alpha or 0.99,
                                             state.epsilon or 1e-8, state.weightDecay or 0))
  M.pack(model)
end

-- Detector:detect in two halves (several frames in flight: one accelerated model / context per frame)
function M.detect_begin(model, img)
  local x = img:float():contiguous()
  model.b200.pending = x                                   -- keep the host frame alive until detect_end
  check(model.b200.ctx, C.frcnn_detect_begin(model.b200.ctx, x:data(), 0, 1, x:size(2), x:size(3)))
end
function M.detect_end(model, out, cap, cnt)
  check(model.b200.ctx, C.frcnn_detect_end(model.b200.ctx, out, cap, cnt))
  model.b200.pending = nil
end
function M.set_schedule(model, throughput)
  check(model.b200.ctx, C.frcnn_set_schedule(model.b200.ctx, throughput and C.FRCNN_SCHED_THROUGHPUT or C.FRCNN_SCHED_LATENCY))
end

return M
