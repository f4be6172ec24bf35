"""GPU twins of tests/test_tf_semantics_pins.py: the CUDA kernels on the same known-answer
fixtures (TensorFlow's published NMS vectors; hand-derived pad-direction / -inf pool / TF1
nearest fixtures), through the C-ABI."""
import numpy as np
import pytest
import torch

from automl_b200 import utils
import test_tf_semantics_pins as pins   # same directory (pytest prepends tests/ to sys.path)

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _ops():
  from automl_b200 import ops
  return ops


def _cuda_nms(boxes, scores, max_out, iou=0.5, score_thresh=float('-inf'), sigma=0.0):
  ops = _ops()
  k = len(scores)
  b = torch.from_numpy(np.asarray(boxes, np.float32)[None]).to(DEV).contiguous()
  s = torch.from_numpy(np.asarray(scores, np.float32)[None]).to(DEV).contiguous()
  c = torch.zeros(1, k, dtype=torch.int32, device=DEV)
  det = torch.empty(1, max_out, 7, device=DEV)
  sel = torch.empty(1, max_out, dtype=torch.int32, device=DEV)
  valid = torch.empty(1, dtype=torch.int32, device=DEV)
  work = torch.empty(ops.nms_work_bytes(1, k), dtype=torch.uint8, device=DEV)
  ops.nms_v5(b, s, c, None, 0, max_out, iou, score_thresh, sigma, (1e6, 1e6), det, sel, valid, work)
  torch.cuda.synchronize()
  v = int(valid[0])
  return sel[0, :v].cpu().numpy().tolist(), det[0, :v, 5].cpu().numpy(), v


def test_cuda_nms_tf_three_clusters():
  assert _cuda_nms(pins.THREE_CLUSTERS, pins.SCORES, 3)[0] == [3, 0, 5]
  assert _cuda_nms(pins.THREE_CLUSTERS_FLIPPED, pins.SCORES, 3)[0] == [3, 0, 5]
  assert _cuda_nms(pins.THREE_CLUSTERS, pins.SCORES, 2)[0] == [3, 0]
  assert _cuda_nms(pins.THREE_CLUSTERS, pins.SCORES, 30)[0] == [3, 0, 5]
  assert _cuda_nms(pins.THREE_CLUSTERS, pins.SCORES - np.float32(5), 6)[0] == [3, 0, 5]
  assert _cuda_nms(pins.THREE_CLUSTERS, pins.SCORES, 3, score_thresh=0.4)[0] == [3, 0]
  ten = np.tile(np.asarray([[0, 0, 1, 1]], np.float32), (10, 1))
  assert _cuda_nms(ten, np.full(10, .9, np.float32), 3)[0] == [0]


def test_cuda_soft_nms_tf_three_clusters():
  idx, sc, valid = _cuda_nms(pins.THREE_CLUSTERS, pins.SCORES, 6, iou=0.5, score_thresh=0.0, sigma=0.5)
  assert idx == [3, 0, 1, 5, 4, 2] and valid == 6
  np.testing.assert_allclose(sc, [0.95, 0.9, 0.384, 0.3, 0.256, 0.197], rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('k,s,size,pad_before', [(3, 2, 8, 0), (3, 2, 7, 1), (5, 2, 8, 1),
                                                  (5, 2, 7, 2), (3, 1, 8, 1), (5, 1, 8, 2)])
def test_cuda_depthwise_same_padding_direction(k, s, size, pad_before):
  """All-ones input, index-coded taps: the corner outputs equal the sum of the taps that fall
  inside the image, which identifies pad_before (and that the extra cell is AFTER)."""
  ops = _ops()
  c = 8
  x = torch.ones(1, size, size, c, dtype=torch.float16, device=DEV)
  taps = torch.arange(k * k, dtype=torch.float32).reshape(k, k) + 1.0
  w = taps.reshape(k * k, 1).expand(k * k, c).contiguous().to(DEV)   # fp32 taps
  n_out = -(-size // s)
  out = torch.empty(1, n_out, n_out, c, dtype=torch.float16, device=DEV)
  ops.depthwise_conv(x, out, w, None, utils.ACT_NONE, k, s)
  torch.cuda.synchronize()
  got = out[0, :, :, 0].float().cpu()
  assert float(got[0, 0]) == float(taps[pad_before:, pad_before:].sum())
  inside = size - ((n_out - 1) * s - pad_before)
  assert float(got[-1, -1]) == float(taps[:inside, :inside].sum())
  assert bool((out[0] == out[0, :, :, :1]).all())      # every channel identical


def test_cuda_max_pool_negative_inputs():
  ops = _ops()
  c = 8
  x = -torch.arange(1, 26, dtype=torch.float32).view(1, 5, 5, 1).expand(1, 5, 5, c).contiguous().half()
  out = torch.empty(1, 3, 3, c, dtype=torch.float16, device=DEV)
  ops.max_pool(x.to(DEV), out, (3, 3), (2, 2))
  torch.cuda.synchronize()
  o = out[0, :, :, 0].float().cpu()
  assert o[2, 2] == -19.0 and o[0, 0] == -1.0 and float(o.max()) < 0
  y = -torch.arange(1, 17, dtype=torch.float32).view(1, 4, 4, 1).expand(1, 4, 4, c).contiguous().half()
  out2 = torch.empty(1, 2, 2, c, dtype=torch.float16, device=DEV)
  ops.max_pool(y.to(DEV), out2, (3, 3), (2, 2))
  torch.cuda.synchronize()
  o2 = out2[0, :, :, 0].float().cpu()
  assert o2[0, 0] == -1.0 and o2[1, 1] == -11.0


@pytest.mark.parametrize('n_in,n_out,expect', [(2, 4, [0, 0, 1, 1]), (3, 5, [0, 0, 1, 1, 2]),
                                               (5, 9, [0, 0, 1, 1, 2, 2, 3, 3, 4]),
                                               (3, 8, [0, 0, 0, 1, 1, 1, 2, 2])])
def test_cuda_fuse_nearest_index_rule(n_in, n_out, expect):
  """fuse_dw with ONE upsampled input, weight 1, no activation and a centre-tap-only depthwise
  kernel returns the resampled map itself: src = min(floor(dst * in/out), in-1)."""
  ops = _ops()
  c = 8
  src = torch.arange(n_in, dtype=torch.float32).view(1, n_in, 1, 1).expand(1, n_in, n_in, c).contiguous().half()
  dwk = torch.zeros(9, c, dtype=torch.float32)
  dwk[4] = 1.0
  out = torch.empty(1, n_out, n_out, c, dtype=torch.float16, device=DEV)
  ops.fuse_dw([(src.to(DEV), ops.RS_UP, None, 1.0)], dwk.to(DEV), out, utils.ACT_NONE)
  torch.cuda.synchronize()
  assert out[0, :, 0, 0].float().cpu().tolist() == [float(v) for v in expect]


def test_cuda_preprocess_half_pixel_bilinear():
  ops = _ops()
  img = np.zeros((1, 2, 2, 3), np.uint8)
  img[:, :, 1, :] = 255
  out = torch.empty(1, 4, 4, 3, dtype=torch.float32, device=DEV)
  scale = ops.preprocess(torch.from_numpy(img).to(DEV), out, [0.0, 0.0, 0.0], [255.0, 255.0, 255.0])
  torch.cuda.synchronize()
  np.testing.assert_allclose(out[0, 0, :, 0].cpu().numpy(), [0.0, 0.25, 0.75, 1.0], atol=1e-6)
  assert scale == pytest.approx(0.5)
