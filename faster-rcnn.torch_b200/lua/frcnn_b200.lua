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
--   b200.accelerate(model)                                          -- plan from the nn modules, binds weights + gradients,
--                                                                   --   overrides pnet/cnet forward/backward (objective.lua,
--                                                                   --   Detector.lua then run unmodified)
--   create_objective = b200.create_objective                        -- optional: the fused per-batch objective
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
    node = node.children and node.children[1]
  end
  return blocks
end

local function dev_ptr(t) return ffi.cast('const float*', t:data()) end
local function dev_ptr_rw(t) return ffi.cast('float*', t:data()) end

-- The reference's model table is { cfg, layers, pnet, cnet } (model_utilities.lua:128-134): the anchor_nets / class_layers
-- tables of models/vgg_*.lua are NOT kept, so they are read back from the nn modules themselves.
--   anchor net i = pnet.outnode.children[i].data.module, an nn.Sequential { SpatialConvolution(k x k), PReLU,
--   SpatialConvolution(1 x 1) } (model_utilities.lua:29-35) whose graph child is the conv block it reads (:51-53);
--   class layers = the first nn.Sequential of cnet: Linear [, BatchNormalization], PReLU [, Dropout] (:80-92).
local function derive_anchor_nets(pnet, blocks)
  local nets, i = {}, 1
  local n_out = #pnet.outnode.children
  for i = 1, n_out - 1 do                                    -- the last output is the conv feature map itself (:55)
    local node = pnet.outnode.children[i]
    local seq = node.data.module
    local conv = seq.modules[1]
    assert(torch.typename(conv) == 'nn.SpatialConvolution' and conv.kW == conv.kH, 'unexpected AnchorNetwork layout')
    local src, input = node.children[1].data.module, nil
    for b, blk in ipairs(blocks) do if blk == src then input = b end end
    assert(input, 'anchor network is not attached to a conv block output')
    nets[i] = { kW = conv.kW, n = conv.nOutputPlane, input = input }
  end
  return nets
end

local function derive_class_layers(cnet)
  local net = cnet:findModules('nn.Sequential')[1]
  local layers = {}
  for _, m in ipairs(net.modules) do
    local t = torch.typename(m)
    if t == 'nn.Linear' then layers[#layers + 1] = { n = m.weight:size(1), dropout = 0, batch_norm = false }
    elseif t == 'nn.BatchNormalization' then layers[#layers].batch_norm = true
    elseif t == 'nn.Dropout' then layers[#layers].dropout = m.p end
  end
  return layers
end

-- Parameter tensors in the library's bind order (frcnn_param_info): trunk convs (weight, bias, PReLU slope), anchor
-- networks (conv weight, bias, slope, 1x1 weight, bias), class layers (Linear weight, bias [, BN weight, bias,
-- running_mean, running_var], slope), Linear(n, 4), Linear(n, classes).  field = 'weight' / 'gradWeight'.
local function walk_params(model, anchor_nets, class_layers, grads)
  local W, B = grads and 'gradWeight' or 'weight', grads and 'gradBias' or 'bias'
  local list, nh = {}, #anchor_nets
  local function push(t) list[#list + 1] = t end
  for _, seq in ipairs(trunk_blocks(model.pnet, nh)) do
    local convs, prelus = seq:findModules('nn.SpatialConvolution'), seq:findModules('nn.PReLU')
    for i = 1, #convs do push(convs[i][W]); push(convs[i][B]); push(prelus[i][W]) end
  end
  for i = 1, nh do
    local seq = model.pnet.outnode.children[i].data.module
    local c1, pr, c2 = seq.modules[1], seq.modules[2], seq.modules[3]
    push(c1[W]); push(c1[B]); push(pr[W]); push(c2[W]); push(c2[B])
  end
  local lin = model.cnet:findModules('nn.Linear')
  local bns = model.cnet:findModules('nn.BatchNormalization')
  local prs = model.cnet:findModules('nn.PReLU')
  local bi = 0
  for i, l in ipairs(class_layers) do
    push(lin[i][W]); push(lin[i][B])
    if l.batch_norm then
      bi = bi + 1
      local bn = bns[bi]
      push(bn[W]); push(bn[B])
      if grads then
        -- running statistics are not parameters in Torch: their gradient slots are never written by the library
        model.b200.bn_dummy = model.b200.bn_dummy or torch.CudaTensor(bn.running_mean:nElement()):zero()
        push(model.b200.bn_dummy); push(model.b200.bn_dummy)
      else
        push(bn.running_mean); push(bn.running_var)
      end
    end
    push(prs[i][W])
  end
  -- the two output branches: Linear(n, 4) and Linear(n, classes) (model_utilities.lua:96-104)
  local reg, cls = lin[#class_layers + 1], lin[#class_layers + 2]
  if reg.weight:size(1) ~= 4 then reg, cls = cls, reg end
  push(reg[W]); push(reg[B]); push(cls[W]); push(cls[B])
  return list
end

-- Builds the plan from the model table alone and binds device pointers of the learnable tensors -- views into the
-- flat CudaTensor created by combine_and_flatten_parameters (utilities.lua:136-147), so optimiser updates (main.lua:133)
-- are seen after the next M.pack(model) -- then installs the module-slot overrides (M.install) so that the reference's
-- objective.lua and Detector.lua run UNMODIFIED on the accelerated model.  Call after load_model (main.lua:114).
function M.accelerate(model, anchor_nets, class_layers)
  local cfg, layers = model.cfg, model.layers
  local n_out = #model.pnet.outnode.children
  anchor_nets = anchor_nets or model.anchor_nets or derive_anchor_nets(model.pnet, trunk_blocks(model.pnet, n_out - 1))
  class_layers = class_layers or model.class_layers or derive_class_layers(model.cnet)
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
  model.b200 = { ctx = ctx, n_heads = nh, anchor_nets = anchor_nets, class_layers = class_layers }
  local tensors = walk_params(model, anchor_nets, class_layers, false)
  assert(#tensors == C.frcnn_param_count(ctx), 'parameter walk does not match the plan')
  local arr = ffi.new('const float*[?]', #tensors)
  for i, t in ipairs(tensors) do arr[i - 1] = dev_ptr(t) end
  check(ctx, C.frcnn_bind_params(ctx, arr, #tensors))
  M.pack(model)
  M.bind_grads(model)
  M.install(model)
  M.install_amp(model)
  return model
end

-- Binds the flat gradient views: the same walk over gradWeight / gradBias (views into the flat `gradient` CudaTensor of
-- combine_and_flatten_parameters, main.lua:92).
function M.bind_grads(model)
  local ctx = model.b200.ctx
  local tensors = walk_params(model, model.b200.anchor_nets, model.b200.class_layers, true)
  local arr = ffi.new('float*[?]', #tensors)
  for i, t in ipairs(tensors) do arr[i - 1] = dev_ptr_rw(t) end
  check(ctx, C.frcnn_bind_grads(ctx, arr, #tensors))
end

-- Module-slot overrides.  nn modules are tables whose methods live in the class metatable, so an instance field shadows
-- the class method: after this, pnet:forward / pnet:backward / cnet:forward / cnet:backward of THIS model run in the
-- library, with the call shapes objective.lua:71,164,179,189 and Detector.lua:33,101 use.  :training() / :evaluate() stay
-- nn.Module's (they set self.train, which selects the mode here); :cuda(), :parameters() are untouched.
function M.install(model)
  local ctx, nh = model.b200.ctx, model.b200.n_heads
  local pnet, cnet = model.pnet, model.cnet
  local dims = ffi.new('int[?]', 3 * (nh + 1))
  local optrs = ffi.new('float*[?]', nh + 1)
  local dptrs = ffi.new('const float*[?]', nh + 1)
  local step = 0

  pnet.forward = function(self, img)                            -- objective.lua:71, Detector.lua:33: img is a 3-D CudaTensor
    local x = img:contiguous()
    local h, w = x:size(2), x:size(3)
    check(ctx, C.frcnn_pnet_output_dims(ctx, h, w, dims))
    self.output = type(self.output) == 'table' and self.output or {}
    for i = 0, nh do
      self.output[i + 1] = self.output[i + 1] or torch.CudaTensor()
      self.output[i + 1]:resize(dims[3 * i], dims[3 * i + 1], dims[3 * i + 2])
      optrs[i] = self.output[i + 1]:data()
    end
    if self.train then
      -- fresh SpatialDropout masks every call: the seed comes from Torch's global generator, as nn.SpatialDropout's does
      step = step + 1
      check(ctx, C.frcnn_pnet_forward_train(ctx, x:data(), 1, h, w, optrs, nil, torch.random()))
      self.b200_input = x                                       -- pnet:backward re-reads the frame (first-layer wgrad)
    else
      check(ctx, C.frcnn_pnet_forward(ctx, x:data(), 1, h, w, optrs))
    end
    return self.output
  end
  pnet.updateOutput = pnet.forward

  pnet.backward = function(self, img, delta_outputs)            -- objective.lua:189; the returned gradInput is unused there
    for i = 0, nh do
      local d = delta_outputs[i + 1]
      dptrs[i] = d and d:contiguous():data() or nil
    end
    check(ctx, C.frcnn_pnet_backward(ctx, dptrs))
    self.gradInput = self.gradInput or torch.CudaTensor()
    return self.gradInput
  end

  cnet.forward = function(self, cinput)                         -- objective.lua:164, Detector.lua:101: R x (kh*kw*C)
    local x = cinput:contiguous()
    local R = x:size(1)
    self.output = type(self.output) == 'table' and self.output or {}
    self.output[1] = (self.output[1] or torch.CudaTensor()):resize(R, 4)
    self.output[2] = (self.output[2] or torch.CudaTensor()):resize(R, model.cfg.class_count + 1)
    if self.train then
      check(ctx, C.frcnn_cnet_forward_train(ctx, x:data(), R, nil, torch.random(), self.output[1]:data(), self.output[2]:data()))
    else
      check(ctx, C.frcnn_cnet_forward(ctx, x:data(), R, self.output[1]:data(), self.output[2]:data()))
    end
    return self.output
  end
  cnet.updateOutput = cnet.forward

  cnet.backward = function(self, cinput, deltas)                -- objective.lua:179: { crdelta, ccdelta } -> post_roi_delta
    self.gradInput = (torch.isTensor(self.gradInput) and self.gradInput or torch.CudaTensor()):resize(cinput:size())
    check(ctx, C.frcnn_cnet_backward(ctx, deltas[1]:contiguous():data(), deltas[2]:contiguous():data(), self.gradInput:data()))
    return self.gradInput
  end
  return model
end

-- The `amp` slot.  objective.lua:30 and Detector.lua:14 build it themselves -- `nn.SpatialAdaptiveMaxPooling(kw, kh):cuda()`
-- is a LOCAL of create_objective / a field set in Detector:__init -- so the constructor is what gets replaced:
-- M.install_amp(model) swaps nn.SpatialAdaptiveMaxPooling for a class with the same surface (forward on the strided crop
-- view extract_roi_pooling_input returns, `indices` readable / assignable, backward -> gradInput of the view's shape)
-- whose updateOutput / updateGradInput are one library launch each.  M.uninstall_amp() puts cunn's class back.
local B200Amp = nil
local function amp_class()
  if B200Amp then return B200Amp end
  B200Amp = torch.class('nn.B200SpatialAdaptiveMaxPooling', 'nn.Module')
  function B200Amp:__init(W, H)
    nn.Module.__init(self)
    self.W, self.H = W, H
    self.indices = torch.Tensor()
  end
  function B200Amp:updateOutput(input)                            -- input: C x h x w view, any strides
    local ctx = B200Amp.ctx
    assert(input:dim() == 3, 'B200 amp: 3-D (C x h x w) input expected (objective.lua:118, Detector.lua:97)')
    local nc, h, w = input:size(1), input:size(2), input:size(3)
    self.output:resize(nc, self.H, self.W)
    self.indices = self.indices:type(input:type()):resize(nc, self.H, self.W)
    check(ctx, C.frcnn_adaptive_maxpool_forward(ctx, input:data(), nc, h, w, input:stride(1), input:stride(2), input:stride(3),
                                                self.H, self.W, self.output:data(), self.indices:data()))
    return self.output
  end
  function B200Amp:updateGradInput(input, gradOutput)
    local ctx = B200Amp.ctx
    local nc, h, w = input:size(1), input:size(2), input:size(3)
    local g = gradOutput:contiguous()
    self.gradInput:resize(nc, h, w)
    check(ctx, C.frcnn_adaptive_maxpool_backward(ctx, g:data(), self.indices:contiguous():data(), nc, h, w, self.H, self.W,
                                                 self.gradInput:data()))
    return self.gradInput
  end
  return B200Amp
end
function M.install_amp(model)
  local cls = amp_class()
  cls.ctx = model.b200.ctx
  M.cunn_amp = M.cunn_amp or nn.SpatialAdaptiveMaxPooling
  nn.SpatialAdaptiveMaxPooling = cls
end
function M.uninstall_amp()
  if M.cunn_amp then nn.SpatialAdaptiveMaxPooling = M.cunn_amp end
end

-- Drop-in for the global create_objective (objective.lua:15): same arguments, same returned closure
-- lossAndGradient(w) -> loss, gradient, same statistics appended to `stats` -- but the per-image loop body
-- (objective.lua:65-198: criteria per anchor, ROI pooling per example, cnet, pnet:backward) is ONE frcnn_train_batch call
-- per group of equally sized frames.  Use: `create_objective = require('frcnn_b200').create_objective` before main.lua:120.
function M.create_objective(model, weights, gradient, batch_iterator, stats)
  local ctx = model.b200.ctx
  local step = 0
  local function cleanAnchors(examples, dims)                   -- objective.lua:32-43
    local i = 1
    while i <= #examples do
      local a = examples[i][1]
      if a.index[2] > dims[3 * (a.layer - 1) + 1] or a.index[3] > dims[3 * (a.layer - 1) + 2] then table.remove(examples, i) else i = i + 1 end
    end
  end
  local dims = ffi.new('int[?]', 3 * (model.b200.n_heads + 1))
  return function(w)
    if w ~= weights then weights:copy(w) end
    M.pack(model)                                               -- the optimiser has moved the weights since the last call
    gradient:zero()
    step = step + 1
    local batch = batch_iterator:nextTraining()
    local groups, order = {}, {}
    for i, x in ipairs(batch) do                                -- frames of one size share a frcnn_train_batch call
      local key = x.img:size(2) .. 'x' .. x.img:size(3)
      if not groups[key] then groups[key] = {}; order[#order + 1] = key end
      table.insert(groups[key], x)
    end
    local cls_loss, reg_loss, creg_loss, ccls_loss = 0, 0, 0, 0
    local cls_count, reg_count, ccls_count = 0, 0, 0
    for _, key in ipairs(order) do
      local g = groups[key]
      local n, h, w_ = #g, g[1].img:size(2), g[1].img:size(3)
      check(ctx, C.frcnn_pnet_output_dims(ctx, h, w_, dims))
      local imgs = torch.CudaTensor(n, 3, h, w_)
      for i, x in ipairs(g) do
        imgs[i]:copy(x.img)                                     -- objective.lua:66: x.img:cuda()
        cleanAnchors(x.positive, dims); cleanAnchors(x.negative, dims)
      end
      local losses = M.train_batch(model, imgs, g, step)
      for i, x in ipairs(g) do
        cls_loss, reg_loss = cls_loss + losses[4 * i - 4], reg_loss + losses[4 * i - 3]
        creg_loss, ccls_loss = creg_loss + losses[4 * i - 2], ccls_loss + losses[4 * i - 1]
        reg_count = reg_count + #x.positive
        cls_count = cls_count + #x.positive + #x.negative
        ccls_count = ccls_count + 1
      end
    end
    gradient:div(cls_count)                                     -- objective.lua:200
    local pcls, preg = cls_loss / cls_count, reg_loss / reg_count
    local dcls, dreg = ccls_loss / ccls_count, creg_loss / reg_count
    print(string.format('prop: cls: %f (%d), reg: %f (%d); det: cls: %f, reg: %f', pcls, cls_count, preg, reg_count, dcls, dreg))
    table.insert(stats.pcls, pcls); table.insert(stats.preg, preg)
    table.insert(stats.dcls, dcls); table.insert(stats.dreg, dreg)
    return pcls + preg, gradient
  end
end

-- frcnn_train_image / frcnn_train_batch return when the losses are on the host; the backward pass may still be adding
-- into the gradient on the context's stream.  cutorch work on the default stream is ordered behind it (the context's
-- stream is a blocking one), so lossAndGradient above needs nothing; a host-side reader (gradient:float(), timing) calls this.
function M.synchronize(model)
  check(model.b200.ctx, C.frcnn_synchronize(model.b200.ctx))
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
                                 seed or torch.random(), losses))   -- fresh masks every call, like nn.SpatialDropout
  return losses[0], losses[1], losses[2], losses[3]
end

-- The same for a group of equally sized frames (frcnn_train_batch): imgs is an n x 3 x h x w CudaTensor, items[i] =
-- { positive = {...}, negative = {...} } as BatchIterator:nextTraining yields them (cleaned).  Dropout seeds are drawn
-- from Torch's global generator.  Returns a 0-based float[4 n] cdata: {cls, reg, creg, ccls} loss sums per frame.
local function fill_example(e, anchor, roi)
  e.anchor[0], e.anchor[1], e.anchor[2], e.anchor[3] = anchor.minX, anchor.minY, anchor.maxX, anchor.maxY
  e.layer, e.aspect, e.y, e.x = anchor.layer, anchor.aspect, anchor.index[2], anchor.index[3]
  if roi then
    local r = roi.rect
    e.roi[0], e.roi[1], e.roi[2], e.roi[3] = r.minX, r.minY, r.maxX, r.maxY
    local t = Anchors.inputToAnchor(anchor, r)            -- FloatTensor(4), objective.lua:110
    for k = 0, 3 do e.reg_target[k] = t[k + 1] end
    e.class_index = roi.class_index
  end
end
function M.train_batch(model, imgs, items, step)
  local ctx, n = model.b200.ctx, #items
  local pos_arr, neg_arr, keep = ffi.new('const frcnn_example*[?]', n), ffi.new('const frcnn_example*[?]', n), {}
  local n_pos, n_neg, seeds = ffi.new('int[?]', n), ffi.new('int[?]', n), ffi.new('uint64_t[?]', n)
  for i, x in ipairs(items) do
    local pos = ffi.new('frcnn_example[?]', math.max(#x.positive, 1))
    local neg = ffi.new('frcnn_example[?]', math.max(#x.negative, 1))
    for k, ex in ipairs(x.positive) do fill_example(pos[k - 1], ex[1], ex[2]) end
    for k, ex in ipairs(x.negative) do fill_example(neg[k - 1], ex[1], nil) end
    keep[#keep + 1] = pos; keep[#keep + 1] = neg               -- keep the cdata alive across the call
    pos_arr[i - 1], neg_arr[i - 1] = pos, neg
    n_pos[i - 1], n_neg[i - 1] = #x.positive, #x.negative
    seeds[i - 1] = torch.random()
  end
  local losses = ffi.new('float[?]', 4 * n)
  check(ctx, C.frcnn_train_batch(ctx, imgs:data(), n, imgs:size(3), imgs:size(4), pos_arr, n_pos, neg_arr, n_neg, nil, seeds, losses))
  return losses
end

-- Data-parallel training inside ONE LuaJIT process (main.lua:52 drives one device; here: one accelerated model per
-- device): models = { model_on_gpu1, model_on_gpu2, ... }.  dp_init forms the NCCL communicator over their contexts,
-- dp_allreduce sums the flat gradients (bound with bind_grads) and the given per-model counter tensors in place.
function M.dp_init(models)
  local n = #models
  local ctxs = ffi.new('frcnn_ctx*[?]', n)
  for i, m in ipairs(models) do ctxs[i - 1] = m.b200.ctx end
  check(models[1].b200.ctx, C.frcnn_dp_init_all(ctxs, n))
end
function M.dp_allreduce(models, counters)       -- counters: optional list of small CudaTensors, one per model
  local n = #models
  local ctxs, ptrs = ffi.new('frcnn_ctx*[?]', n), ffi.new('float*[?]', n)
  for i, m in ipairs(models) do
    ctxs[i - 1] = m.b200.ctx
    ptrs[i - 1] = counters and counters[i]:data() or nil
  end
  check(models[1].b200.ctx, C.frcnn_dp_allreduce(ctxs, n, counters and ptrs or nil, counters and counters[1]:nElement() or 0))
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

-- The nearby-aversion candidates of BatchIterator.lua:206-217 (before shuffle_n) as one launch: positives =
-- { {anchor_rect, roi}, ... }; returns { {anchor_rect}, ... } in the reference's order.
function M.find_nearby_negative(model, anchors, positives, neg_threshold)
  local ctx, n = model.b200.ctx, #positives
  if n == 0 then return {} end
  local pos = ffi.new('frcnn_anchor_ref[?]', n)
  for i, p in ipairs(positives) do
    local a = p[1]
    pos[i - 1].layer, pos[i - 1].aspect, pos[i - 1].y, pos[i - 1].x = a.layer, a.aspect, a.index[2], a.index[3]
  end
  local cap = 64 * n
  local out, out_pos, cnt = ffi.new('frcnn_anchor_ref[?]', cap), ffi.new('int[?]', cap), ffi.new('int[1]')
  check(ctx, C.frcnn_find_nearby_negative(ctx, pos, n, neg_threshold, out, out_pos, cap, cnt))
  local found = {}
  for i = 0, cnt[0] - 1 do found[i + 1] = { anchors:get(out[i].layer, out[i].aspect, out[i].y, out[i].x) } end
  return found
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
