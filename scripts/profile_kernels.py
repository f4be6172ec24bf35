"""Launches representative kernels of the D0@640 batch-32 step stand-alone (for `ncu --set full`).
Usage: python scripts/profile_kernels.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automl_b200 import ops, utils  # noqa: E402

dev = 'cuda:0'
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = 32


def dw(h, c, k, s, se=True):
  x = torch.randn(N, h, h, c, device=dev).half()
  w = torch.randn(k * k, c, device=dev)
  b = torch.randn(c, device=dev)
  ho = -(-h // s)
  out = torch.empty(N, ho, ho, c, dtype=torch.float16, device=dev)
  part = torch.zeros(N, c, dtype=torch.int64, device=dev) if se else None
  for _ in range(reps):
    ops.depthwise_conv(x, out, w, b, utils.ACT_SWISH, k, s, part)


def pw(m, k, n, act, res=False, per_image=False):
  a = torch.randn(N, m // N, k, device=dev).half()
  w = (torch.randn(N if per_image else 1, n, k, device=dev) / k**0.5).half()
  b = torch.randn(n, device=dev)
  out = torch.empty(N, m // N, (n + 7) // 8 * 8, dtype=torch.float16, device=dev)
  r = torch.randn_like(out) if res else None
  for _ in range(reps):
    ops.pointwise_conv(a, w, b, out, act, residual=r, rows=m // N, batch=N, nout=n)


def stem():
  x = torch.randn(N, 640, 640, 3, device=dev)
  w = torch.randn(27, 32, device=dev).half()
  b = torch.randn(32, device=dev)
  out = torch.empty(N, 320, 320, 32, dtype=torch.float16, device=dev)
  for _ in range(reps):
    ops.stem_conv(x, out, w, b, utils.ACT_SWISH)


def front(h, cin, cmid, k, s):
  x = torch.randn(N, h, h, cin, device=dev).half()
  we = (torch.randn(cmid, cin, device=dev) / cin**0.5).half()
  be = torch.randn(cmid, device=dev) * 0.1
  wd = torch.randn(k * k, cmid, device=dev) / k
  bd = torch.randn(cmid, device=dev) * 0.1
  ho = -(-h // s)
  out = torch.empty(N, ho, ho, cmid, dtype=torch.float16, device=dev)
  se = torch.zeros(N, cmid, dtype=torch.int64, device=dev)
  for _ in range(reps):
    ops.mbconv_expand_dw(x, we, be, wd, bd, out, utils.ACT_SWISH, k, s, se)


def tower(h, f=64):
  x = torch.randn(N, h, h, f, device=dev).half()
  dwk = torch.randn(9, f, device=dev) / 3
  pwk = (torch.randn(f, f, device=dev) / f**0.5).half()
  b = torch.randn(f, device=dev) * 0.1
  out = torch.empty(N, h, h, f, dtype=torch.float16, device=dev)
  for _ in range(reps):
    ops.sepconv([(x, ops.RS_SAME, None, 1.0)], utils.ACT_NONE, dwk, pwk, b, out, utils.ACT_SWISH)


def node(h, f=64):
  a = torch.randn(N, h, h, f, device=dev).half()
  u = torch.randn(N, h // 2, h // 2, f, device=dev).half()
  dwk = torch.randn(9, f, device=dev) / 3
  out = torch.empty(N, h, h, f, dtype=torch.float16, device=dev)
  for _ in range(reps):
    ops.fuse_dw([(a, ops.RS_SAME, None, 0.6), (u, ops.RS_UP, None, 0.4)], dwk, out, utils.ACT_SWISH)


stem()
front(320, 16, 96, 3, 2)   # blocks_1 fused expand + dw
tower(80)                  # head tower layer, level 3
node(80)                   # BiFPN td node, level 3
dw(320, 32, 3, 1)          # blocks_0 dw
dw(320, 96, 3, 2)          # blocks_1 dw (largest byte mover)
dw(80, 240, 5, 1)          # blocks_4 dw
dw(40, 672, 5, 1)          # blocks_9 dw
pw(3276800, 16, 96, utils.ACT_SWISH)                 # blocks_1 expand
pw(819200, 144, 24, utils.ACT_NONE, res=True, per_image=True)   # blocks_2 project
pw(51200, 80, 480, utils.ACT_SWISH)                  # blocks_6 expand
pw(12800, 1152, 192, utils.ACT_NONE, res=True, per_image=True)  # blocks_12 project
pw(204800, 64, 810, utils.ACT_NONE)                  # class-predict L3 (forward(): logits stored)


def class_argmax(h, f=64, a=9, c=90):
  x = torch.randn(N, h, h, f, device=dev).half()
  w = torch.zeros(a, ops.CLASS_ARGMAX_COLS, f, device=dev)
  w[:, :c] = torch.randn(a, c, f, device=dev) / f**0.5
  b = torch.full((a, ops.CLASS_ARGMAX_COLS), float('-inf'), device=dev)
  b[:, :c] = -4.6
  total = h * h * a
  scores = torch.empty(N, total, device=dev)
  classes = torch.empty(N, total, dtype=torch.int32, device=dev)
  wp, bp = w.reshape(-1, f).half().contiguous(), b.reshape(-1).contiguous()
  for _ in range(reps):
    ops.class_argmax(x, wp, bp, scores, classes, 0, a)


class_argmax(80)           # class head fused with the class arg-max, level 3 (detect path)
torch.cuda.synchronize()
print('done')
