"""Times the fused expand+depthwise kernel against pointwise + depthwise at the D0 (640, batch 32)
block shapes.  CUDA events, 20 iterations after 3 warm-ups; prints one JSON line per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from automl_b200 import ops, utils

DEV = 'cuda:0'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SHAPES = [  # name, h, cin, cmid, k, s
    ('b1', 320, 16, 96, 3, 2), ('b2', 160, 24, 144, 3, 1), ('b3', 160, 24, 144, 5, 2),
    ('b4', 80, 40, 240, 5, 1), ('b5', 80, 40, 240, 3, 2), ('b6', 40, 80, 480, 3, 1),
    ('b8', 40, 80, 480, 5, 1), ('b9', 40, 112, 672, 5, 1), ('b11', 40, 112, 672, 5, 2),
    ('b12', 20, 192, 1152, 5, 1),
]


def timeit(fn, iters=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters


ONLY = sys.argv[2].split(',') if len(sys.argv) > 2 else None
for name, h, cin, cmid, k, s in SHAPES:
  if ONLY and name not in ONLY:
    continue
  g = torch.Generator().manual_seed(1)
  x = torch.randn(N, h, h, cin, generator=g).half().to(DEV)
  we = (torch.randn(cmid, cin, generator=g) / cin**0.5).half().to(DEV)
  be = (torch.randn(cmid, generator=g) * 0.2).to(DEV)
  wk = (torch.randn(k * k, cmid, generator=g) / k).to(DEV)
  bd = (torch.randn(cmid, generator=g) * 0.1).to(DEV)
  ho = -(-h // s)
  e = torch.empty(N, h, h, cmid, dtype=torch.float16, device=DEV)
  o1 = torch.empty(N, ho, ho, cmid, dtype=torch.float16, device=DEV)
  o2 = torch.empty_like(o1)
  se1 = torch.zeros(N, cmid, dtype=torch.int64, device=DEV)
  se2 = torch.zeros_like(se1)

  def unfused():
    ops.pointwise_conv(x, we, be, e, utils.ACT_SWISH, rows=h * h, batch=N)
    ops.depthwise_conv(e, o1, wk, bd, utils.ACT_SWISH, k, s, se1)

  def fused():
    ops.mbconv_expand_dw(x, we, be, wk, bd, o2, utils.ACT_SWISH, k, s, se2)

  tu, tf = timeit(unfused), timeit(fused)
  se1.zero_(); se2.zero_()
  unfused(); fused()
  torch.cuda.synchronize()
  err = float((o1.float() - o2.float()).abs().max())
  se_err = float((se1 - se2).abs().max()) / 2.0**20
  gb_f = 2.0 * N * (h * h * cin + ho * ho * cmid) / 1e9
  print(json.dumps({'block': name, 'unfused_ms': round(tu, 4), 'fused_ms': round(tf, 4),
                    'fused_GBps': round(gb_f / (tf * 1e-3), 1), 'max_abs_diff': err,
                    'se_abs_diff': se_err}))
