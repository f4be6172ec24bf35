"""Prints the NMS fast-path outcome per image and times the NMS launches on D0@640."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automl_b200 import arch, hparams_config, weights
from automl_b200.engine import Engine

c = hparams_config.get_efficientdet_config('efficientdet-d0'); c.override(dict(image_size=640))
a = arch.DetArch(c); w = weights.synthetic_weights(a, 0)
n = 32
eng = Engine(c, w, n, use_cuda_graph=False)
x = np.random.default_rng(0).uniform(0, 1, size=(n, 640, 640, 3)).astype(np.float32)
eng.detect(torch.from_numpy(x)); torch.cuda.synchronize()
flags = eng._post[eng._cur]['work'][-4 * n:].view(torch.int32).cpu().numpy()
print('fast-path reason codes per image (0 = fast path proved exact):', flags.tolist())
print('valid:', eng.valid.cpu().numpy().tolist())
nms = [fn for name, fn in eng._ops if name == 'nms'][0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
  e0.record(); nms(); e1.record(); torch.cuda.synchronize()
  print('nms (fast + full launch) %.3f ms' % e0.elapsed_time(e1))
