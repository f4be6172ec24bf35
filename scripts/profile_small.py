"""Representative small / mid kernels of the D0@640 batch-32 step (for ncu --set full)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automl_b200 import ops, utils
dev = 'cuda:0'; N = 32
# stem
x = torch.rand(N, 640, 640, 3, device=dev)
w = torch.randn(27, 32, device=dev).half(); b = torch.randn(32, device=dev)
out = torch.empty(N, 320, 320, 32, dtype=torch.float16, device=dev)
ops.stem_conv(x, out, w, b, utils.ACT_SWISH)
# BiFPN node at P3 (80x80x64): same + upsampled input
F = 64
same = torch.randn(N, 80, 80, F, device=dev).half(); up = torch.randn(N, 40, 40, F, device=dev).half()
dwk = torch.randn(9, F, device=dev).half(); o2 = torch.empty(N, 80, 80, F, dtype=torch.float16, device=dev)
ops.fuse_dw([(same, ops.RS_SAME, None, 0.6), (up, ops.RS_UP, None, 0.4)], dwk, o2, utils.ACT_SWISH)
# BiFPN / head pointwise 64 -> 64 at P3 and P4
for rows in (204800, 51200, 12800):
  a = torch.randn(rows, F, device=dev).half(); wt = torch.randn(F, F, device=dev).half(); bb = torch.randn(F, device=dev)
  oo = torch.empty(rows, F, dtype=torch.float16, device=dev)
  ops.pointwise_conv(a, wt, bb, oo, utils.ACT_SWISH, rows=rows, batch=1)
# head depthwise 3x3 at P3
hx = torch.randn(N, 80, 80, F, device=dev).half(); ho = torch.empty_like(hx)
ops.depthwise_conv(hx, ho, dwk, None, utils.ACT_NONE, 3, 1)
torch.cuda.synchronize(); print('done')
