import sys, torch
sys.path.insert(0, "/root/repo")
import torch.nn.functional as TF
import frcnn_b200 as F
from oracle import model as OM
case = tuple(int(v) for v in sys.argv[1:8])
n, h, w, cin, cout, k, pad = case
m = F.vgg_small(F.duplo_cfg)
g = torch.Generator().manual_seed(1)
x = OM.bf16_round(torch.randn(n, cin, h, w, generator=g)).requires_grad_(True)
wt = OM.bf16_round(torch.randn(cout, cin, k, k, generator=g) * 0.05).requires_grad_(True)
y = TF.conv2d(x, wt, None, padding=pad)
dy = OM.bf16_round(torch.randn(y.shape, generator=g))
y.backward(dy)
x_nhwc = x.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
dw = torch.zeros(cout, cin, k, k).cuda()
ffi, L = F.ffi, F.lib()
rc = L.frcnn_conv_wgrad_bf16(m.ctx, ffi.cast("const uint16_t*", x_nhwc.data_ptr()), ffi.cast("const uint16_t*", dy_nhwc.data_ptr()),
                             n, h, w, cin, cout, k, pad, ffi.cast("float*", dw.data_ptr()))
if rc != 0:
    print(case, "rc", rc, ffi.string(L.frcnn_last_error(m.ctx))[:100]); sys.exit(0)
err = (dw.cpu() - wt.grad).abs().max().item()
print(case, "ok max err", err, "ref max", wt.grad.abs().max().item())
