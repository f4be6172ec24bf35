"""world_size-2 gloo test of the batch-shard + all-gather host logic (runs on CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from automl_b200 import parallel


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, batch, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  lo, hi = parallel.shard_range(rank, world, batch)
  # what each rank's NMS kernel writes: rows [image_id, ymin, xmin, ymax, xmax, score, class]
  local = torch.zeros(batch, 100, 7)
  local[:, :, 0] = torch.arange(lo, hi, dtype=torch.float32)[:, None]
  local[:, :, 5] = float(rank + 1)
  out = parallel.gather_detections(local)
  dist.barrier()
  if rank == 0:
    q.put(out.numpy())
  dist.destroy_process_group()


def test_shard_range():
  assert parallel.shard_range(0, 8, 32) == (0, 32)
  assert parallel.shard_range(7, 8, 32) == (224, 256)
  with pytest.raises(ValueError):
    parallel.shard_range(8, 8, 32)


def test_gather_detections_two_ranks():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, q)) for r in range(2)]
  for p in procs:
    p.start()
  out = q.get(timeout=120)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert out.shape == (6, 100, 7)
  np.testing.assert_array_equal(out[:, 0, 0], np.arange(6, dtype=np.float32))   # global image order
  np.testing.assert_array_equal(out[:, 0, 5], [1, 1, 1, 2, 2, 2])


def test_gather_single_process_is_identity():
  t = torch.zeros(2, 100, 7)
  assert parallel.gather_detections(t) is t
