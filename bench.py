"""EfficientDet-D0 images/sec at batch 32 per GPU on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (stem .. heads .. pre-NMS .. NMS) over one batch of 32
synthetic 640x640 images per GPU.  `value` is whole-job images/s with inputs already resident in
HBM; `e2e` is the same metric through the public call with HOST buffers (pinned host images
copied H2D and the [B,100,7] detections copied D2H inside the timed region).  Inputs
(157 MB fp32 per batch) are larger than the 126 MB L2, and the activations written between
kernels (GBs per step) flush it, so no explicit L2 flush is needed between iterations.

--impl reference times the reference's CPU implementation of the path.  TensorFlow is not
installable offline, so it is the oracle port (oracle/efficientdet_oracle.py +
oracle/postprocess_oracle.py) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

MODEL = 'efficientdet-d0'
IMAGE_SIZE = 640
BATCH = 32
METRIC = 'EfficientDet-D0 images/sec @ batch 32/GPU (640x640, forward + post-process)'


def build_config():
  from automl_b200 import hparams_config
  c = hparams_config.get_efficientdet_config(MODEL)
  c.override(dict(image_size=IMAGE_SIZE))
  return c


class ClockSampler(object):
  """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
  Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index=0):
    self.index, self.samples, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
           '--format=csv,noheader,nounits', '-lms', '100'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      threading.Thread(target=self._read, daemon=True).start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.samples.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    sm, mx, reasons = [], [], set()
    for s in self.samples:
      parts = [p.strip() for p in s.split(',')]
      if len(parts) < 6:
        continue
      try:
        sm.append(float(parts[0])); mx.append(float(parts[1]))
      except ValueError:
        continue
      for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                            'sw_power_cap'), parts[2:6]):
        if val.lower().startswith('active'):
          reasons.add(name)
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return d['hbm_gbs'], d.get('bf16_tflops_sustained', d.get('bf16_tflops')), 'measured'
  return 6650.0, 1400.0, 'fallback'


def cpu_baseline(config, weights, images, seconds_budget=25.0):
  """Oracle port on the host cores: bounded sample of the same workload."""
  import numpy as np
  import torch
  from oracle import efficientdet_oracle as eo
  from oracle import postprocess_oracle as po
  cores = min(os.cpu_count() or 1, 32)   # more threads than this slow the small convs down
  torch.set_num_threads(cores)
  orc = eo.Oracle(config, weights, torch.float32)
  params = config.as_dict()
  sample = images[:1]
  t0 = time.perf_counter()
  cls_o, box_o = orc(sample)
  po.det_post_process(params, {l: v.numpy() for l, v in cls_o.items()},
                      {l: v.numpy() for l, v in box_o.items()}, np.ones(1, np.float32))
  one = time.perf_counter() - t0
  nimg = int(max(1, min(BATCH, len(images), seconds_budget // max(one, 1e-3))))
  sample = images[:nimg]
  t0 = time.perf_counter()
  cls_o, box_o = orc(sample)
  po.det_post_process(params, {l: v.numpy() for l, v in cls_o.items()},
                      {l: v.numpy() for l, v in box_o.items()}, np.ones(nimg, np.float32))
  dt = time.perf_counter() - t0
  return {'value': nimg / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
          'sample': '%d image(s) of the %dx%d batch through the oracle port (PyTorch-CPU network '
                    '+ numpy post-process), %.1f s' % (nimg, IMAGE_SIZE, IMAGE_SIZE, dt)}


def workload_config(world):
  """The `config` object of the JSON line: identical for our arm and the reference arm."""
  return {'workload': 'EfficientDet-D0 640x640 batch %d/GPU: stem, 16 MBConv, 3 BiFPN cells, '
                      'class/box heads, pre-NMS, NMS-V5 (gaussian)' % BATCH,
          'global_batch': world * BATCH, 'parallelism': 'batch-shard x%d' % world,
          'l2': 'inputs (157 MB fp32) and per-step activations (GBs) exceed the 126 MB L2, '
                'so every timed iteration starts with a flushed L2'}


def run_reference(args, rank, world):
  """--impl reference: the reference's CPU implementation of the path (oracle port)."""
  if rank != 0:
    return
  import numpy as np
  from automl_b200 import arch, weights as weights_lib
  config = build_config()
  w = weights_lib.synthetic_weights(arch.DetArch(config), 0)
  x = np.random.default_rng(0).uniform(0, 1, size=(8, IMAGE_SIZE, IMAGE_SIZE, 3)).astype(np.float32)
  per_step_budget = max(2.0, 120.0 / max(1, args.steps + args.warmup))
  vals = []
  base = None
  for i in range(args.warmup + args.steps):
    base = cpu_baseline(config, w, x, seconds_budget=per_step_budget)
    if i >= args.warmup:
      vals.append(base['value'])
  v = float(sum(vals) / len(vals))
  base['value'] = v
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * BATCH / v,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic',
      'config': dict(workload_config(world),
                     note='TensorFlow is not installable offline: the oracle port of the reference '
                          'path on the host cores, a bounded sample of the batch per step'),
      'cpu_baseline': base,
      'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--profile-out', default='')
  args = ap.parse_args()

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if args.impl == 'reference':
    run_reference(args, rank, world)
    return

  import numpy as np
  import torch
  import torch.distributed as dist
  import __graft_entry__
  __graft_entry__.build()
  from automl_b200 import arch, weights as weights_lib
  from automl_b200.engine import Engine

  torch.cuda.set_device(local_rank)
  dev = 'cuda:%d' % local_rank
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(dev))

  config = build_config()
  from automl_b200 import inference, parallel
  rng = np.random.default_rng(rank)
  # COCO-shaped raw input: 480x640 uint8 images in pinned host memory
  host_raw = torch.from_numpy(
      rng.integers(0, 256, size=(BATCH, 480, 640, 3), dtype=np.uint8)).pin_memory()
  driver = inference.ServingDriver(MODEL, '_', batch_size=BATCH,
                                   model_params={'image_size': IMAGE_SIZE}, device=dev,
                                   image_id_base=rank * BATCH)
  driver.build()
  eng = driver.engine
  w = None
  gathered = torch.empty(world * BATCH, eng.max_output_size, 7, device=dev) if world > 1 else None

  gather_hook = (lambda det: parallel.gather_detections(det, gathered)) if world > 1 else None

  def step(e2e):
    if e2e:
      # the public call: host uint8 images in, host detections out (H2D + D2H + sync inside)
      return driver.serve_images(host_raw)
    # network + pre-NMS on the main stream; NMS (+ the single collective of the path, the
    # all-gather of per-image detections) on the engine's NMS stream, overlapping the next step
    eng.run(postprocess=True, after_nms=gather_hook)
    return None

  def timed(e2e, steps, warmup):
    for _ in range(warmup):
      step(e2e)
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      step(e2e)
    if not e2e:
      eng.wait_detections()      # the last step's NMS / all-gather is inside the timed region
    e1.record()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  driver.serve_images(host_raw)   # builds the graph and leaves a pre-processed batch in HBM
  torch.cuda.synchronize()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  ms_dev = timed(False, args.steps, max(3, args.warmup))
  clocks = sampler.stop() if rank == 0 else None
  ms_e2e = timed(True, args.steps, 3)

  value = world * BATCH * args.steps / (ms_dev / 1000.0)
  e2e_value = world * BATCH * args.steps / (ms_e2e / 1000.0)

  line = None
  if rank == 0:
    hbm_peak, tf_peak, peak_src = measured_peaks()
    rows = eng.profile_ops(iters=3)
    kinds = {}
    for r in rows:
      k = kinds.setdefault(r['kind'], {'ms': 0.0, 'bytes': 0, 'flops': 0, 'launches': 0})
      k['ms'] += r['ms']; k['bytes'] += r['bytes']; k['flops'] += r['flops']; k['launches'] += 1
    total_ms = sum(k['ms'] for k in kinds.values())
    dom = max(kinds, key=lambda n: kinds[n]['ms'])
    # depthwise kernels of all (k, stride) instantiations are one kernel family
    dw = {'ms': 0.0, 'bytes': 0, 'launches': 0}
    for name, k in kinds.items():
      if name.startswith('depthwise'):
        dw['ms'] += k['ms']; dw['bytes'] += k['bytes']; dw['launches'] += k['launches']
    pw = kinds.get('pointwise_tc', {'ms': 0.0, 'bytes': 0, 'flops': 0, 'launches': 0})
    if dw['ms'] >= pw['ms']:
      fam, fam_name = dw, 'depthwise_kernel (all k/stride)'
    else:
      fam, fam_name = pw, 'pointwise_tc_kernel'
    achieved = fam['bytes'] / (fam['ms'] / 1e3) / 1e9
    roofline = {'kernel': fam_name, 'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak,
                'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': None,
                'peak_source': peak_src, 'share_of_step': fam['ms'] / total_ms,
                'launches': fam['launches'],
                'per_kind': {n: {'ms': round(k['ms'], 4), 'GBps': round(k['bytes'] / max(k['ms'], 1e-9) / 1e6, 1),
                                 'TFLOPs': round(k['flops'] / max(k['ms'], 1e-9) / 1e9, 2),
                                 'launches': k['launches']} for n, k in sorted(kinds.items())}}
    # DRAM bytes of the dominant kernel family per forward, from the committed ncu capture of
    # this same command (profiles/r1_traffic.json, made by scripts/gpu_final.sh)
    tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    if os.path.exists(tpath):
      with open(tpath) as f:
        fam_key = 'depthwise_kernel' if fam is dw else 'pointwise_tc_kernel'
        t = json.load(f)['per_forward'].get(fam_key)
      if t:
        roofline['traffic'] = int((t['dram_read_MB'] + t['dram_write_MB']) * 1e6)
        roofline['traffic_source'] = 'profiles/r1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, family sum per forward)'
    if args.profile_out:
      with open(args.profile_out, 'w') as f:
        json.dump({'ops': rows, 'kinds': roofline['per_kind'], 'sum_ms': total_ms}, f, indent=1)
    base = None
    if not args.no_cpu_baseline:
      w = weights_lib.synthetic_weights(arch.DetArch(config), 0)
      base = cpu_baseline(config, w, eng.input.cpu().numpy())
    line = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f16 storage / f32 accumulate', 'data': 'synthetic',
        'config': workload_config(world),
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': int(host_raw.numel()) + 4 * BATCH,
                'd2h_bytes_per_step': int(world * BATCH * eng.max_output_size * 7 * 4),
                'api': 'inference.ServingDriver.serve_images(uint8 [32,480,640,3] pinned host) -> '
                       'numpy detections (H2D, device pre-process, network, NMS, all-gather, D2H, sync)'},
        'gpu_launches': eng.launches_per_forward * args.steps,
        'nms_full_queue_images': eng.nms_fallback_count(),
        'clocks': clocks, 'roofline': roofline, 'cpu_baseline': base,
    }
    print(json.dumps(line))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
