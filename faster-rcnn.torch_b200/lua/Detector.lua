-- Drop-in replacement of the reference's Detector.lua: same class, same constructor, same detect() result table
-- ({p, a, r, l, r2, class, confidence} per winner); the body of detect() is ONE call into libfrcnn_b200.so.
-- Not runnable in this repository's image (no Lua); see frcnn_b200.lua.
require 'Anchors'
local ffi = require 'ffi'
local b200 = require 'frcnn_b200'

local Detector = torch.class('Detector')

function Detector:__init(model)
  assert(model.b200, 'call require("frcnn_b200").accelerate(model) first')
  self.model = model
  self.anchors = Anchors.new(model.pnet, model.cfg.scales)       -- unchanged Lua class (host-side LUTs)
  self.localizer = Localizer.new(model.pnet.outnode.children[5])
  self.cap = 4096
  self.out = ffi.new('frcnn_detection[?]', self.cap)
  self.n = ffi.new('int[1]')
end

function Detector:detect(input)
  local ctx = self.model.b200.ctx
  local img = input:float():contiguous()                         -- host CHW FloatTensor, copied to the GPU inside
  b200.check(ctx, b200.C.frcnn_detect(ctx, img:data(), 1, img:size(2), img:size(3), self.out, self.cap, self.n))
  local winners = {}
  for i = 0, self.n[0] - 1 do
    local d = self.out[i]
    table.insert(winners, {
      p = d.p, l = d.layer, class = d.cls, confidence = d.confidence,
      a = self.anchors:get(d.layer, d.aspect, d.y, d.x),
      r = Rect.new(d.r[0], d.r[1], d.r[2], d.r[3]),
      r2 = Rect.new(d.r2[0], d.r2[1], d.r2[2], d.r2[3]),
    })
  end
  return winners
end
