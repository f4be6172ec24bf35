"""Batch-shard data parallelism of the serving path: one process per GPU, weights replicated,
rank r owns images [r*B, (r+1)*B) of the global batch, and the ONLY collective is one all-gather
of the per-image detection blocks after NMS (the reference's closest
analogue is tf2/eval.py:64-113, per-replica model + NMS under MirroredStrategy).

`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the plumbing.
"""
import torch
import torch.distributed as dist


def shard_range(rank, world_size, per_rank_batch):
  """(first global image index, one-past-last) owned by `rank`; also the image_id_base."""
  if not 0 <= rank < world_size:
    raise ValueError('rank %d outside world of %d' % (rank, world_size))
  return rank * per_rank_batch, (rank + 1) * per_rank_batch


def gather_detections(local, out=None, group=None):
  """All-gathers [B, max_out, 7] detection blocks into [world*B, max_out, 7] (rank order ==
  global image order, because rank r wrote image ids r*B .. (r+1)*B-1)."""
  if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
    return local
  world = dist.get_world_size(group)
  if out is None:
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
  dist.all_gather_into_tensor(out, local.contiguous(), group=group)
  return out
