"""Where the graph-replayed step goes: replays prefixes of the launch list (backbone, + feature
network, + heads, + pre-NMS) as CUDA graphs and reports the marginal time of each segment.
usage: python scripts/time_segments.py [d0|d4|d7x] [out.json]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from automl_b200 import arch, weights  # noqa: E402
from automl_b200.engine import Engine  # noqa: E402


def main():
  name = sys.argv[1] if len(sys.argv) > 1 else 'd0'
  cfg = bench.CONFIGS[name]
  c = bench.build_config(cfg)
  a = arch.DetArch(c)
  eng = Engine(c, weights.synthetic_weights(a, 0), cfg['batch'])
  s = cfg['image_size']
  eng.set_input(torch.from_numpy(np.random.default_rng(0).uniform(-2, 2, size=(cfg['batch'], s, s, 3)).astype(np.float32)))
  names = eng.op_names()
  last_block = max(i for i, n in enumerate(names) if n.startswith('blocks_'))
  first_head = min(i for i, n in enumerate(names) if n.startswith('class_net') or n.startswith('box_net'))
  cuts = [('stem..blocks', last_block + 1), ('+ feature network', first_head),
          ('+ heads', eng.num_network_ops), ('+ pre-NMS', len(names) - 1)]
  # the launches of run(postprocess=True): class head fused with the class arg-max
  eng._fused_target = eng._post[0] if eng.fuse_class_argmax else None  # pylint: disable=protected-access
  out, prev = [], 0.0
  for label, upto in cuts:
    fn = lambda upto=upto: eng._run_ops(upto)  # pylint: disable=protected-access
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      fn()
    for _ in range(5):
      g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
      g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out.append({'prefix': label, 'launches': upto, 'ms': round(ms, 4), 'marginal_ms': round(ms - prev, 4)})
    prev = ms
    print(json.dumps(out[-1]))
  if len(sys.argv) > 2:
    with open(sys.argv[2], 'w') as f:
      json.dump(out, f, indent=1)


if __name__ == '__main__':
  main()
