"""Exploration: D4 / D7x at a small image size vs the oracle (prints per-tensor rel-L2)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automl_b200 import arch, hparams_config, weights  # noqa: E402
from automl_b200.engine import Engine  # noqa: E402
from oracle import efficientdet_oracle as eo  # noqa: E402


def rel_l2(a, b):
  a, b = a.double().flatten(), b.double().flatten()
  return float((a - b).norm() / max(float(b.norm()), 1e-30))


for name, size in (('efficientdet-d4', 256), ('efficientdet-d7x', 256)):
  c = hparams_config.get_efficientdet_config(name)
  c.override(dict(image_size=size))
  a = arch.DetArch(c)
  w = weights.synthetic_weights(a, 0)
  x = np.random.default_rng(1).uniform(-2, 2, size=(1, size, size, 3)).astype(np.float32)
  t0 = time.time()
  orc = eo.Oracle(c, w, torch.float32)
  cls_ref, box_ref = orc(x)
  t1 = time.time()
  eng = Engine(c, w, 1, use_cuda_graph=False)
  cls_out, box_out = eng.forward(torch.from_numpy(x))
  torch.cuda.synchronize()
  bb = [rel_l2(eng.buffers[b.name + '/out'].float().cpu().permute(0, 3, 1, 2), orc.endpoints[b.name]) for b in a.blocks]
  fp = [rel_l2(eng.fpn_feats[l].float().cpu().permute(0, 3, 1, 2), orc.endpoints['fpn_%d' % l]) for l in a.levels]
  ce = [rel_l2(cls_out[l].float().cpu(), cls_ref[l]) for l in a.levels]
  be = [rel_l2(box_out[l].float().cpu(), box_ref[l]) for l in a.levels]
  amax = max(float(eng.buffers[b.name + '/out'].float().abs().max()) for b in a.blocks)
  print(name, 'oracle %.1fs' % (t1 - t0), 'blocks', len(a.blocks), 'levels', a.levels, 'F', a.fpn_filters,
        'fusion', a.fpn_weight_method)
  print('  backbone worst %.2e (last %.2e) max|act| %.1f' % (max(bb), bb[-1], amax))
  print('  fpn', ['%.2e' % v for v in fp])
  print('  cls', ['%.2e' % v for v in ce])
  print('  box', ['%.2e' % v for v in be])
