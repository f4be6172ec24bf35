"""CPU oracle for the EfficientNet V1/V2 backbone: the reference's
efficientnetv2/effnetv2_model.py forward pass restated in PyTorch (CPU, fp32 / fp64).

TEST INFRASTRUCTURE ONLY (same rule as efficientdet_oracle.py): nothing under automl_b200/
imports this module; tests/ use it as the checker.

Parity status: "parity unpinned" for the convolution numerics (TensorFlow is not installable
here, Appendix C of SURVEY.md); the structure is pinned by the reference's RNG-free goldens:
`count_params()` of 15 models (effnetv2_model_test.py:25-48) and the block / filter rounding
rules.  Functions cite /root/reference/efficientnetv2/effnetv2_model.py lines.
"""
import torch

from oracle import efficientdet_oracle as eo


class EffNetV2Oracle(object):
  """call(images NHWC float) -> dict of endpoints (NCHW tensors): 'stem', 'block_i',
  'reduction_i', 'features', 'head_1x1' (EffNetV2Model.call :595-658)."""

  def __init__(self, arch, weights, dtype=torch.float32, store=None):
    self.arch = arch
    self.dtype = dtype
    self.w = {k: torch.as_tensor(v).to(dtype) for k, v in weights.items()}
    self.store = store or (lambda t: t)     # e.g. eo.fp16_store to model fp16 activations
    self.act = lambda t: eo.activation_fn(t, arch.mconfig.act_fn)

  def _bn(self, x, scope):
    return eo.batch_norm_inference(x, self.w, scope, self.arch.bn_eps)

  def _se(self, x, sc):
    """SE.call :135-147: reduce_mean -> conv(+bias) -> act -> conv(+bias) -> sigmoid * x."""
    w = self.w
    s = x.mean((2, 3), keepdim=True)
    s = eo.conv2d_same(s, w[sc + '/se/conv2d/kernel']) + w[sc + '/se/conv2d/bias'].view(1, -1, 1, 1)
    s = self.act(s)
    s = eo.conv2d_same(s, w[sc + '/se/conv2d_1/kernel']) + w[sc + '/se/conv2d_1/bias'].view(1, -1, 1, 1)
    return torch.sigmoid(s) * x

  def _block(self, b, x):
    w, sc = self.w, '%s/%s' % (self.arch.model_name, b.name)
    convs = iter(['conv2d', 'conv2d_1'])
    bns = iter(['tpu_batch_normalization', 'tpu_batch_normalization_1', 'tpu_batch_normalization_2'])
    inputs = x
    if b.conv_type == 0:      # MBConvBlock.call :279-311
      if b.expand_ratio != 1:
        x = self.store(self.act(self._bn(eo.conv2d_same(x, w['%s/%s/kernel' % (sc, next(convs))]),
                                         '%s/%s' % (sc, next(bns)))))
      x = self.act(self._bn(eo.depthwise_conv2d_same(x, w[sc + '/depthwise_conv2d/depthwise_kernel'],
                                                     b.strides), '%s/%s' % (sc, next(bns))))
      if b.se_filters:
        # the device folds the gate into the project weights: the depthwise output is stored,
        # the gated tensor is not
        x = self._se(self.store(x), sc)
      else:
        x = self.store(x)
      x = self._bn(eo.conv2d_same(x, w['%s/%s/kernel' % (sc, next(convs))]), '%s/%s' % (sc, next(bns)))
    else:                     # FusedMBConvBlock.call :375-406
      if b.expand_ratio != 1:
        x = self.store(self.act(self._bn(
            eo.conv2d_same(x, w['%s/%s/kernel' % (sc, next(convs))], b.strides), '%s/%s' % (sc, next(bns)))))
      if b.se_filters:
        x = self._se(x, sc)
      stride = 1 if b.expand_ratio != 1 else b.strides
      x = self._bn(eo.conv2d_same(x, w['%s/%s/kernel' % (sc, next(convs))], stride), '%s/%s' % (sc, next(bns)))
      if b.expand_ratio == 1:
        x = self.act(x)       # add act if no expansion (:401-402)
    if b.has_skip:            # residual :270-277 (drop_connect is the identity at inference)
      x = x + inputs
    return self.store(x)

  def __call__(self, images):
    a, w, mn = self.arch, self.w, self.arch.model_name
    x = torch.as_tensor(images).to(self.dtype).permute(0, 3, 1, 2)
    ep = {}
    x = self.store(self.act(self._bn(eo.conv2d_same(x, w[mn + '/stem/conv2d/kernel'], 2),
                                     mn + '/stem/batch_normalization')))      # Stem :409-432
    ep['stem'] = x
    red = 0
    for i, b in enumerate(a.blocks):
      x = self._block(b, x)
      ep['block_%d' % i] = x
      if i in a.reductions:
        red += 1
        ep['reduction_%d' % red] = x
    ep['features'] = x
    x = self.store(self.act(self._bn(eo.conv2d_same(x, w[mn + '/head/conv2d/kernel']),
                                     mn + '/head/batch_normalization')))      # Head :472-474
    ep['head_1x1'] = x
    return ep
