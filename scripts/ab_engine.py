"""A/B timing of Engine variants inside ONE process (same box, same clocks): D0 640 batch 32,
graph replay with pipelined NMS (what bench.py reports as `value`).
usage: ab_engine.py name=kw:val,kw:val name2=...   e.g.  base=fuse_mbconv_front:0 fused=fuse_mbconv_front:1"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automl_b200 import hparams_config, weights as weights_lib  # noqa: E402
from automl_b200.arch import DetArch  # noqa: E402
from automl_b200.engine import Engine  # noqa: E402

BATCH, SIZE = 32, 640
cfg = hparams_config.get_detection_config('efficientdet-d0')
cfg.image_size = SIZE
cfg.is_training_bn = False
w = weights_lib.synthetic_weights(DetArch(cfg), 0)
variants = []
for spec in sys.argv[1:]:
  name, _, kws = spec.partition('=')
  kw = {}
  for item in filter(None, kws.split(',')):
    k, _, v = item.partition(':')
    kw[k] = int(v)
  variants.append((name, Engine(cfg, w, BATCH, **kw)))
x = torch.from_numpy(np.random.default_rng(0).standard_normal((BATCH, SIZE, SIZE, 3)).astype(np.float32)).cuda()
for _, e in variants:
  e.set_input(x)


POST = os.environ.get('AB_POSTPROCESS', '1') == '1'   # AB_POSTPROCESS=0: network only (no pre-NMS / NMS)


def timed(eng, steps=30, warmup=5):
  for _ in range(warmup):
    eng.run(postprocess=POST)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(steps):
    eng.run(postprocess=POST)
  if POST:
    eng.wait_detections()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / steps


for rep in range(3):
  print(' '.join('%s=%.3fms' % (n, timed(e)) for n, e in variants), flush=True)
