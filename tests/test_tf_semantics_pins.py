"""Pins for the THIRD-PARTY (TensorFlow) op semantics the oracle restates.

TensorFlow is not installable here, so these are (a) the known-answer vectors of TensorFlow's own
published unit tests for NonMaxSuppression V2-V5, quoted with their source, and (b) hand-derived
fixtures -- small enough to verify with pencil and paper -- that DISTINGUISH the possible
conventions (pad-right vs pad-left for 'SAME' stride-2 windows, -inf vs zero padding in max-pool,
floor vs round in the TF1 nearest-neighbour resize, half-pixel centres in bilinear resize).
The `-m gpu` twins of the same fixtures run the CUDA kernels (tests/test_gpu_kernels.py).
"""
import numpy as np
import pytest
import torch

from oracle import efficientdet_oracle as eo
from oracle import postprocess_oracle as po

# ---- tensorflow/core/kernels/image/non_max_suppression_op_test.cc and
# ---- tensorflow/python/ops/image_ops_test.py (NonMaxSuppressionWithScoresTest) ------------------
THREE_CLUSTERS = np.asarray([[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9], [0, 10, 1, 11],
                             [0, 10.1, 1, 11.1], [0, 100, 1, 101]], np.float32)
THREE_CLUSTERS_FLIPPED = np.asarray([[1, 1, 0, 0], [0, 0.1, 1, 1.1], [0, .9, 1, -0.1],
                                     [0, 10, 1, 11], [1, 10.1, 0, 11.1], [1, 101, 0, 100]], np.float32)
SCORES = np.asarray([.9, .75, .6, .95, .5, .3], np.float32)
NEG_INF = float('-inf')


def _nms(boxes, scores, max_out, iou=0.5, score_thresh=NEG_INF, sigma=0.0, pad=False):
  return po.non_max_suppression_v5(boxes, scores, max_out, iou, score_thresh, sigma, pad)


def test_tf_select_from_three_clusters():                       # TestSelectFromThreeClusters
  idx, sc, valid = _nms(THREE_CLUSTERS, SCORES, 3)
  assert idx.tolist() == [3, 0, 5] and valid == 3
  np.testing.assert_array_equal(sc, np.asarray([.95, .9, .3], np.float32))


def test_tf_flipped_coordinates():              # TestSelectFromThreeClustersFlippedCoordinates
  idx, _, _ = _nms(THREE_CLUSTERS_FLIPPED, SCORES, 3)
  assert idx.tolist() == [3, 0, 5]


def test_tf_at_most_two_boxes():                # TestSelectAtMostTwoBoxesFromThreeClusters
  assert _nms(THREE_CLUSTERS, SCORES, 2)[0].tolist() == [3, 0]


def test_tf_at_most_thirty_boxes():             # TestSelectAtMostThirtyBoxesFromThreeClusters
  idx, _, valid = _nms(THREE_CLUSTERS, SCORES, 30)
  assert idx.tolist() == [3, 0, 5] and valid == 3


def test_tf_negative_scores():                  # TestSelectWithNegativeScores
  assert _nms(THREE_CLUSTERS, SCORES - np.float32(5), 6)[0].tolist() == [3, 0, 5]


def test_tf_score_threshold():                  # V3: TestSelectFromThreeClustersWithScoreThreshold
  assert _nms(THREE_CLUSTERS, SCORES, 3, score_thresh=0.4)[0].tolist() == [3, 0]


def test_tf_single_box_and_identical_boxes():   # TestSelectSingleBox, TestSelectFromTenIdenticalBoxes
  assert _nms(np.asarray([[0, 0, 1, 1]], np.float32), np.asarray([.9], np.float32), 3)[0].tolist() == [0]
  ten = np.tile(np.asarray([[0, 0, 1, 1]], np.float32), (10, 1))
  assert _nms(ten, np.full(10, .9, np.float32), 3)[0].tolist() == [0]


def test_tf_zero_max_output_and_empty_input():  # TestSelectFromThreeClustersWithZeroOutput.., TestEmptyInput
  assert _nms(THREE_CLUSTERS, SCORES, 0)[2] == 0
  assert _nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 30)[2] == 0


def test_tf_soft_nms_three_clusters():
  """image_ops_test.py NonMaxSuppressionWithScoresTest.testSelectFromThreeClustersWithSoftNMS:
  max_output_size 6, iou_threshold 0.5, score_threshold 0.0, soft_nms_sigma 0.5 ->
  indices [3, 0, 1, 5, 4, 2], scores [0.95, 0.9, 0.384, 0.3, 0.256, 0.197] (rtol / atol 1e-2).
  Box 1 overlaps box 0 with IoU 0.818 > 0.5 and still survives with a decayed score: the IoU
  threshold does not hard-suppress in soft mode; 0.197 = 0.6 * exp(-0.818^2) * exp(-0.667^2) shows
  the decay is applied once per selected box, newest first."""
  idx, sc, valid = _nms(THREE_CLUSTERS, SCORES, 6, iou=0.5, score_thresh=0.0, sigma=0.5)
  assert idx.tolist() == [3, 0, 1, 5, 4, 2] and valid == 6
  np.testing.assert_allclose(sc, [0.95, 0.9, 0.384, 0.3, 0.256, 0.197], rtol=1e-2, atol=1e-2)
  iou01 = 0.9 / 1.1
  iou12 = 0.8 / 1.2
  np.testing.assert_allclose(sc[2], 0.75 * np.exp(-iou01**2), rtol=1e-6)
  np.testing.assert_allclose(sc[5], 0.6 * np.exp(-iou01**2) * np.exp(-iou12**2), rtol=1e-6)


def test_tf_pad_to_max_output_size():
  """V4/V5 pad_to_max_output_size: indices and scores are zero padded, valid_outputs counts the
  real ones (postprocess.py:193-205 relies on this)."""
  idx, sc, valid = _nms(THREE_CLUSTERS, SCORES, 5, pad=True)
  assert idx.tolist() == [3, 0, 5, 0, 0] and valid == 3
  np.testing.assert_array_equal(sc, np.asarray([.95, .9, .3, 0, 0], np.float32))


# ---- 'SAME' padding direction (tf.nn.convolution docs: pad_before = total // 2, extra AFTER) ------
def _impulse_response(k, s, size):
  """depthwise conv of an index-coded kernel over an all-ones [1,1,size,size] map: the output
  at (0, 0) is the sum of the kernel taps that fall INSIDE the image, which identifies how many
  rows / columns of padding sit before the first element."""
  x = torch.ones(1, 1, size, size, dtype=torch.float64)
  w = torch.arange(k * k, dtype=torch.float64).reshape(k, k, 1, 1) + 1.0
  return eo.depthwise_conv2d_same(x, w, s)[0, 0]


@pytest.mark.parametrize('k,s,size,pad_before', [
    (3, 2, 8, 0),    # even size, k3 s2: total pad 1 -> (0, 1): nothing before (pad-left would give 1)
    (3, 2, 7, 1),    # odd size: total 2 -> (1, 1)
    (5, 2, 8, 1),    # total 3 -> (1, 2)
    (5, 2, 7, 2),    # total 4 -> (2, 2)
    (3, 1, 8, 1), (5, 1, 8, 2),
])
def test_same_padding_puts_the_extra_cell_after(k, s, size, pad_before):
  out = _impulse_response(k, s, size)
  taps = torch.arange(k * k, dtype=torch.float64).reshape(k, k) + 1.0
  # top-left output: kernel rows / cols [pad_before:] overlap the image
  assert float(out[0, 0]) == float(taps[pad_before:, pad_before:].sum())
  n_out = -(-size // s)
  last_start = (n_out - 1) * s - pad_before          # first input row under the last window
  inside = size - last_start                          # rows of the window inside the image
  assert float(out[-1, -1]) == float(taps[:inside, :inside].sum())
  assert eo.same_pad_amounts(size, k, s)[0] == pad_before


def test_stride2_conv_on_even_size_reads_rows_0_1_2_first():
  """A k3 s2 'SAME' conv on an even-sized map: output (0,0) must read input rows/cols 0..2
  (pad (0,1)); a pad-left convention would read -1..1.  Coded input: value = 10*y + x."""
  x = (10.0 * torch.arange(6).view(6, 1) + torch.arange(6).view(1, 6)).double().view(1, 1, 6, 6)
  w = torch.zeros(3, 3, 1, 1, dtype=torch.float64)
  w[2, 2] = 1.0                                       # picks the bottom-right tap
  out = eo.depthwise_conv2d_same(x, w, 2)[0, 0]
  assert out[0, 0] == 22.0 and out[1, 1] == 44.0 and out[2, 2] == 0.0   # (4+2, 4+2) is padding


# ---- max-pool SAME: padded cells never win (-inf), even over negative inputs ------------------
def test_max_pool_same_negative_inputs():
  x = -torch.arange(1, 26, dtype=torch.float32).view(1, 1, 5, 5)    # all negative
  out = eo.max_pool_same(x, (3, 3), (2, 2))[0, 0]
  assert out.shape == (3, 3)
  # window of output (2,2) = rows/cols 3..5 -> only (3..4, 3..4) exist: max = x[3,3] = -19
  assert out[2, 2] == -19.0 and out[0, 0] == -1.0
  assert float(out.max()) < 0                       # zero padding would have produced 0 somewhere
  # even size: pad (0,1) -> output (0,0) covers rows/cols 0..2
  y = -torch.arange(1, 17, dtype=torch.float32).view(1, 1, 4, 4)
  o2 = eo.max_pool_same(y, (3, 3), (2, 2))[0, 0]
  assert o2.shape == (2, 2) and o2[0, 0] == -1.0 and o2[1, 1] == -11.0


# ---- TF1 nearest-neighbour resize: src = min(floor(dst * in/out), in - 1) -----------------------
@pytest.mark.parametrize('n_in,n_out,expect', [
    (2, 4, [0, 0, 1, 1]),                     # exact 2x: dst // 2
    (3, 5, [0, 0, 1, 1, 2]),                  # floor(dst * 0.6): 0, .6, 1.2, 1.8, 2.4
    (5, 9, [0, 0, 1, 1, 2, 2, 3, 3, 4]),      # P-level pair of an odd-sized pyramid (9 -> 5 -> 9)
    (4, 7, [0, 0, 1, 1, 2, 2, 3]),
    (3, 8, [0, 0, 0, 1, 1, 1, 2, 2]),         # half-pixel centres would give [0,0,0,1,1,2,2,2]
])
def test_tf1_nearest_index_rule(n_in, n_out, expect):
  x = torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in, 1).expand(1, 1, n_in, n_in).contiguous()
  out = eo.resize_nearest_tf1(x, n_out, n_out)[0, 0, :, 0]
  assert out.tolist() == [float(v) for v in expect]


# ---- serving pre-process bilinear (tf.image.resize v2: half-pixel centres, no antialias) ------
def test_bilinear_half_pixel_centres_2x():
  """Upscaling [0, 1] by 2 with half-pixel centres gives [0, .25, .75, 1] (align-corners would
  give [0, 1/3, 2/3, 1]; the TF1 asymmetric rule [0, .5, 1, 1])."""
  img = np.zeros((2, 2, 3), np.float32)
  img[:, 1, :] = 255.0
  out, scale = po.image_preprocess(img.astype(np.uint8), 4, [0.0, 0.0, 0.0], [255.0, 255.0, 255.0])
  np.testing.assert_allclose(out[0, :, 0], [0.0, 0.25, 0.75, 1.0], atol=1e-6)
  assert scale == pytest.approx(0.5)
