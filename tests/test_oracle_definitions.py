"""Definition-level pins of the oracle's TensorFlow-owned ops.

TensorFlow cannot run here, so its floating-point RESULTS cannot be recorded; what can be done is
to hold the oracle's torch-based restatements to an INDEPENDENT implementation of TensorFlow's
documented definitions, written as plain numpy loops from the formulas in the TF API reference
(tf.nn.convolution 'SAME' padding: out = ceil(in / stride), pad_total = max((out - 1) * stride +
k - in, 0), pad_before = pad_total // 2; tf.nn.depthwise_conv2d; tf.nn.max_pool2d with 'SAME' where
padded cells never win; tf.compat.v1.image.resize_nearest_neighbor with align_corners=False;
inference batch normalisation (x - mean) * gamma / sqrt(var + eps) + beta).  Random tensors, odd and
even sizes, strides 1 and 2, float64 on both sides -> agreement to 1e-12.
"""
import numpy as np
import pytest
import torch

from oracle import efficientdet_oracle as eo


def _same_pads(size, k, s):
  out = -(-size // s)
  total = max((out - 1) * s + k - size, 0)
  return out, total // 2


def _conv_def(x, w, s):
  """x [H, W, Cin], w [kh, kw, Cin, Cout] -> [Ho, Wo, Cout]; zero padding, extra cell AFTER."""
  h, wd, cin = x.shape
  kh, kw, _, cout = w.shape
  ho, pt = _same_pads(h, kh, s)
  wo, pl = _same_pads(wd, kw, s)
  out = np.zeros((ho, wo, cout))
  for y in range(ho):
    for xx in range(wo):
      for ky in range(kh):
        for kx in range(kw):
          iy, ix = y * s + ky - pt, xx * s + kx - pl
          if 0 <= iy < h and 0 <= ix < wd:
            out[y, xx] += x[iy, ix] @ w[ky, kx]
  return out


def _depthwise_def(x, w, s):
  h, wd, c = x.shape
  kh, kw, _, _ = w.shape
  ho, pt = _same_pads(h, kh, s)
  wo, pl = _same_pads(wd, kw, s)
  out = np.zeros((ho, wo, c))
  for y in range(ho):
    for xx in range(wo):
      for ky in range(kh):
        for kx in range(kw):
          iy, ix = y * s + ky - pt, xx * s + kx - pl
          if 0 <= iy < h and 0 <= ix < wd:
            out[y, xx] += x[iy, ix] * w[ky, kx, :, 0]
  return out


def _max_pool_def(x, k, s):
  h, wd, c = x.shape
  ho, pt = _same_pads(h, k, s)
  wo, pl = _same_pads(wd, k, s)
  out = np.full((ho, wo, c), -np.inf)
  for y in range(ho):
    for xx in range(wo):
      for ky in range(k):
        for kx in range(k):
          iy, ix = y * s + ky - pt, xx * s + kx - pl
          if 0 <= iy < h and 0 <= ix < wd:
            out[y, xx] = np.maximum(out[y, xx], x[iy, ix])
  return out


def _nchw(a):
  return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)[None]))


def _hwc(t):
  return t[0].permute(1, 2, 0).numpy()


@pytest.mark.parametrize('h,w,k,s', [(7, 9, 3, 1), (8, 8, 3, 2), (9, 7, 3, 2), (10, 11, 5, 1),
                                     (12, 9, 5, 2), (5, 5, 1, 1), (1, 1, 3, 2)])
def test_conv_and_depthwise_same_match_the_definition(h, w, k, s):
  rng = np.random.default_rng(h * 100 + w * 10 + k + s)
  x = rng.normal(size=(h, w, 6))
  wk = rng.normal(size=(k, k, 6, 4))
  got = _hwc(eo.conv2d_same(_nchw(x), torch.from_numpy(wk), s))
  np.testing.assert_allclose(got, _conv_def(x, wk, s), rtol=1e-12, atol=1e-12)
  wd = rng.normal(size=(k, k, 6, 1))
  got = _hwc(eo.depthwise_conv2d_same(_nchw(x), torch.from_numpy(wd), s))
  np.testing.assert_allclose(got, _depthwise_def(x, wd, s), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('h,w,k,s', [(8, 8, 3, 2), (9, 7, 3, 2), (5, 4, 2, 1), (7, 7, 3, 1), (1, 1, 3, 2)])
def test_max_pool_same_matches_the_definition(h, w, k, s):
  rng = np.random.default_rng(h + 31 * w + k)
  x = rng.normal(size=(h, w, 5)) - 3.0          # mostly negative: zero padding would win
  got = _hwc(eo.max_pool_same(_nchw(x), (k, k), (s, s)))
  np.testing.assert_array_equal(got, _max_pool_def(x, k, s))


@pytest.mark.parametrize('n_in,n_out', [(3, 8), (5, 10), (4, 7), (10, 20), (1, 5)])
def test_nearest_resize_matches_the_definition(n_in, n_out):
  rng = np.random.default_rng(n_in * 7 + n_out)
  x = rng.normal(size=(n_in, n_in + 1, 3))
  got = _hwc(eo.resize_nearest_tf1(_nchw(x), n_out, 2 * n_out))
  ys = np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * (np.float32(n_in) / np.float32(n_out))),
                  n_in - 1).astype(int)
  xs = np.minimum(np.floor(np.arange(2 * n_out, dtype=np.float32) *
                           (np.float32(n_in + 1) / np.float32(2 * n_out))), n_in).astype(int)
  np.testing.assert_array_equal(got, x[ys][:, xs])


def test_batch_norm_inference_matches_the_definition():
  rng = np.random.default_rng(4)
  x = rng.normal(size=(6, 5, 7))
  w = {'s/gamma': rng.uniform(0.5, 1.5, 7), 's/beta': rng.normal(size=7),
       's/moving_mean': rng.normal(size=7), 's/moving_variance': rng.uniform(0.2, 2.0, 7)}
  wt = {k: torch.from_numpy(v) for k, v in w.items()}
  got = _hwc(eo.batch_norm_inference(_nchw(x), wt, 's', 1e-3))
  want = (x - w['s/moving_mean']) * w['s/gamma'] / np.sqrt(w['s/moving_variance'] + 1e-3) + w['s/beta']
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
