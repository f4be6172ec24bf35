"""A/B of the two depthwise implementations on the D0 @ 640 batch-32 shapes (and the head / D4
shapes): CUDA-event time per launch with a >126 MB L2 flush between launches.
usage: python scripts/ab_depthwise.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automl_b200 import ops, utils  # noqa: E402

DEV = 'cuda:0'
SHAPES = [  # name, n, h, w, c, k, s, se
    ('d0 blocks_0', 32, 320, 320, 32, 3, 1, True), ('d0 blocks_1', 32, 320, 320, 96, 3, 2, True),
    ('d0 blocks_2', 32, 160, 160, 144, 3, 1, True), ('d0 blocks_3', 32, 160, 160, 144, 5, 2, True),
    ('d0 blocks_4', 32, 80, 80, 240, 5, 1, True), ('d0 blocks_5', 32, 80, 80, 240, 3, 2, True),
    ('d0 blocks_6', 32, 40, 40, 480, 3, 1, True), ('d0 blocks_8', 32, 40, 40, 480, 5, 1, True),
    ('d0 blocks_9', 32, 40, 40, 672, 5, 1, True), ('d0 blocks_11', 32, 40, 40, 672, 5, 2, True),
    ('d0 blocks_12', 32, 20, 20, 1152, 5, 1, True), ('d0 head l3', 32, 80, 80, 64, 3, 1, False),
    ('d4 blocks_4', 8, 256, 256, 192, 3, 1, True), ('d4 blocks_10', 8, 128, 128, 336, 5, 1, True),
]


def main():
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
  rows = []
  for name, n, h, w, c, k, s, se in SHAPES:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, h, w, c, generator=g).half().to(DEV)
    wk = (torch.randn(k * k, c, generator=g) / k).to(DEV)
    bias = (torch.randn(c, generator=g) * 0.1).to(DEV) if se else None
    ho, wo = -(-h // s), -(-w // s)
    out = torch.empty(n, ho, wo, c, dtype=torch.float16, device=DEV)
    part = torch.zeros(n, c, dtype=torch.int64, device=DEV) if se else None
    act = utils.ACT_SWISH if se else utils.ACT_NONE
    nbytes = 2 * n * c * (h * w + ho * wo)
    row = {'name': name, 'shape': [n, h, w, c], 'k': k, 's': s, 'MB': nbytes / 1e6}
    for impl, tag in ((1, 'register'), (0, 'tiled')):
      ops.set_option('dw_impl', impl)
      ts = []
      for it in range(7):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.depthwise_conv(x, out, wk, bias, act, k, s, part)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
      ts = sorted(ts[2:])
      ms = ts[len(ts) // 2]
      row[tag + '_us'] = round(ms * 1e3, 1)
      row[tag + '_GBps'] = round(nbytes / ms / 1e6, 0)
    ops.set_option('dw_impl', 0)
    rows.append(row)
    print(json.dumps(row))
  if len(sys.argv) > 1:
    with open(sys.argv[1], 'w') as f:
      json.dump(rows, f, indent=1)


if __name__ == '__main__':
  main()
