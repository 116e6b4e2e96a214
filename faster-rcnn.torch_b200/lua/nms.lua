-- Drop-in replacement of the reference's nms.lua: same global, same arguments, same return value
-- (LongTensor of 1-based indices in pick order), computed by the CUDA kernels of libfrcnn_b200.so.
-- Not runnable in this repository's image (no Lua); see frcnn_b200.lua.
local ffi = require 'ffi'
local b200 = require 'frcnn_b200'
local ctx

function nms(boxes, overlap, scores)
  local pick = torch.LongTensor()
  if boxes:numel() == 0 then return pick end          -- nms.lua:26-28
  -- nms.lua:37-43: number -> that column, 'area' -> area, ANYTHING else (including a score tensor) -> max-y
  local mode, col = 0, 0
  if type(scores) == 'number' then mode, col = 2, scores - 1 elseif scores == 'area' then mode = 1 end
  ctx = ctx or b200.create()
  local b = boxes:float():contiguous()                 -- the reference runs nms on CPU FloatTensors (Detector.lua:74)
  local n = b:size(1)
  pick:resize(n)
  local cnt = ffi.new('int64_t[1]')
  b200.check(ctx, b200.C.frcnn_nms(ctx, b:data(), n, b:size(2), overlap, mode, col, ffi.cast('int64_t*', pick:data()), cnt))
  local k = tonumber(cnt[0])
  if k == 0 then return torch.LongTensor() end
  return pick:narrow(1, 1, k):add(1)                   -- 0-based -> Lua's 1-based
end
