"""CPU oracle: the reference's EfficientDet forward restated in PyTorch (CPU, fp32/fp64).

TEST INFRASTRUCTURE ONLY. Nothing under automl_b200/ may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
use it, and only as the checker / reported CPU baseline.

Parity status: the network arithmetic lives in TensorFlow (pinned >=2.10,<2.16,
/root/reference/efficientdet/requirements.txt:8), which is NOT installed here, so the
conv / pool / resize semantics below are restated from TF's documented behaviour
 and pinned only through the reference's RNG-free tests:
parameter counts for D0-D7 (efficientdet_arch_test.py:47-90), backbone endpoint shapes
(:159-167), fuse_features numerics (:187-236), BiFPN node lists
(tf2/fpn_configs_test.py:24-57), feat sizes / activations (utils_test.py:67-127).
=> "parity unpinned" for the convolution numerics themselves (no TF golden tensors can
be generated offline); see DESIGN.md.

Each function cites the reference file:line it follows (paths under
/root/reference/efficientdet).
"""
import collections

import numpy as np
import torch
import torch.nn.functional as F

from oracle import structure_oracle

# Inference BatchNorm epsilon of the EfficientNet backbones and of the detector's own layers
# (backbone/efficientnet_builder.py:191, efficientnet_lite_builder.py:69; utils.py:244-257).
BN_EPSILON = 1e-3


# ------------------------------------------------------------------------------------
# TF op semantics (third-party)
# ------------------------------------------------------------------------------------
def same_pad_amounts(in_size, k, s):
  """TF 'SAME': out = ceil(in/s); extra padding goes to bottom/right."""
  out = -(-in_size // s)
  total = max((out - 1) * s + k - in_size, 0)
  return total // 2, total - total // 2


def conv2d_same(x, w_hwio, stride=1, groups=1):
  """tf.keras.layers.Conv2D(padding='same', use_bias=False) on NCHW x; w is HWIO."""
  kh, kw = w_hwio.shape[0], w_hwio.shape[1]
  pt, pb = same_pad_amounts(x.shape[2], kh, stride)
  pl, pr = same_pad_amounts(x.shape[3], kw, stride)
  x = F.pad(x, (pl, pr, pt, pb))
  w = w_hwio.permute(3, 2, 0, 1).contiguous()  # OIHW
  return F.conv2d(x, w, stride=stride, groups=groups)


def depthwise_conv2d_same(x, w_hwc1, stride=1):
  """tf.keras.layers.DepthwiseConv2D(padding='same'); kernel [kh,kw,C,1]."""
  kh, kw, c, _ = w_hwc1.shape
  pt, pb = same_pad_amounts(x.shape[2], kh, stride)
  pl, pr = same_pad_amounts(x.shape[3], kw, stride)
  x = F.pad(x, (pl, pr, pt, pb))
  w = w_hwc1.permute(2, 3, 0, 1).contiguous()  # [C,1,kh,kw]
  return F.conv2d(x, w, stride=stride, groups=c)


def max_pool_same(x, pool, stride):
  """tf.layers.max_pooling2d(padding='SAME'): padded cells never win (-inf)."""
  pt, pb = same_pad_amounts(x.shape[2], pool[0], stride[0])
  pl, pr = same_pad_amounts(x.shape[3], pool[1], stride[1])
  x = F.pad(x, (pl, pr, pt, pb), value=float('-inf'))
  return F.max_pool2d(x, kernel_size=pool, stride=stride)


def resize_nearest_tf1(x, out_h, out_w):
  """tf.image.resize_nearest_neighbor (align_corners=False, half_pixel_centers=False):
  src = min(floor(dst * (in/out)), in-1), scale computed in float32."""
  in_h, in_w = x.shape[2], x.shape[3]
  hs = np.float32(in_h) / np.float32(out_h)
  ws = np.float32(in_w) / np.float32(out_w)
  iy = np.minimum(np.floor(np.arange(out_h, dtype=np.float32) * hs),
                  in_h - 1).astype(np.int64)
  ix = np.minimum(np.floor(np.arange(out_w, dtype=np.float32) * ws),
                  in_w - 1).astype(np.int64)
  return x[:, :, torch.from_numpy(iy)][:, :, :, torch.from_numpy(ix)]


def batch_norm_inference(x, w, scope, eps):
  """utils.py:244-326 / util_keras.py:29-66 at is_training=False."""
  g, b = w[scope + '/gamma'], w[scope + '/beta']
  m, v = w[scope + '/moving_mean'], w[scope + '/moving_variance']
  scale = g / torch.sqrt(v + eps)
  return x * scale.view(1, -1, 1, 1) + (b - m * scale).view(1, -1, 1, 1)


def activation_fn(x, act_type):
  """utils.py:36-53."""
  if act_type in ('silu', 'swish', 'swish_native'):
    return x * torch.sigmoid(x)
  if act_type == 'hswish':
    return x * F.relu6(x + 3) / 6
  if act_type == 'relu':
    return F.relu(x)
  if act_type == 'relu6':
    return F.relu6(x)
  if act_type == 'mish':
    return x * torch.tanh(F.softplus(x))
  raise ValueError('Unsupported act_type {}'.format(act_type))


# ------------------------------------------------------------------------------------
# Network
# ------------------------------------------------------------------------------------
class Oracle(object):
  """Forward pass of efficientdet_arch.efficientdet on CPU.

  Args:
    config: hparams_config.Config (already overridden).
    weights: dict reference-variable-name -> numpy array (Keras layouts).
    dtype: torch.float32 or torch.float64.
    store: optional callable applied to every tensor that the CUDA engine writes to
      HBM (used to model fp16 storage rounding); identity by default.
  """

  def __init__(self, config, weights, dtype=torch.float32, store=None):
    self.p = config
    self.dtype = dtype
    self.w = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in weights.items()}
    self.store = store or (lambda t: t)
    self.endpoints = collections.OrderedDict()

  # -- backbone: efficientnet_model.py:710-779, MBConvBlock :372-414, SE :183-195 -----
  def backbone(self, x):
    p = self.p
    name = p.backbone_name
    # the oracle's own walk of the block strings (oracle/structure_oracle.py), NOT the product's
    _, blocks = structure_oracle.backbone_blocks(name)
    eps = BN_EPSILON
    act = lambda t: activation_fn(t, p.act_type)
    w = self.w
    # Stem :526-527
    x = conv2d_same(x, w[name + '/stem/conv2d/kernel'], stride=2)
    x = self.store(act(batch_norm_inference(x, w, name + '/stem/tpu_batch_normalization', eps)))
    self.endpoints['stem'] = x
    feats = {}
    for b in blocks:
      scope = '%s/%s' % (name, b['name'])
      inputs = x
      if b['expand_conv']:
        x = conv2d_same(x, w['%s/%s/kernel' % (scope, b['expand_conv'])])
        x = self.store(act(batch_norm_inference(x, w, '%s/%s' % (scope, b['expand_bn']), eps)))
      x = depthwise_conv2d_same(x, w[scope + '/depthwise_conv2d/depthwise_kernel'], b['stride'])
      x = self.store(act(batch_norm_inference(x, w, '%s/%s' % (scope, b['dw_bn']), eps)))
      if b['has_se']:
        se = x.mean(dim=(2, 3), keepdim=True)
        se = conv2d_same(se, w[scope + '/se/conv2d/kernel']) + w[scope + '/se/conv2d/bias'].view(1, -1, 1, 1)
        se = act(se)
        se = conv2d_same(se, w[scope + '/se/conv2d_1/kernel']) + w[scope + '/se/conv2d_1/bias'].view(1, -1, 1, 1)
        x = torch.sigmoid(se) * x
      x = conv2d_same(x, w['%s/%s/kernel' % (scope, b['project_conv'])])
      x = batch_norm_inference(x, w, '%s/%s' % (scope, b['project_bn']), eps)
      if b['has_skip']:
        x = x + inputs
      x = self.store(x)
      self.endpoints[b['name']] = x
      if b['reduction']:
        feats[b['reduction']] = x
    return feats, eps

  # -- efficientdet_arch.py:55-132 --------------------------------------------------------
  def resample_feature_map(self, feat, scope, target_h, target_w, eps):
    p = self.p
    w = self.w
    _, c, h, wd = feat.shape

    def maybe_1x1(t):
      if c != p.fpn_num_filters:
        t = conv2d_same(t, w[scope + '/conv2d/kernel']) + w[scope + '/conv2d/bias'].view(1, -1, 1, 1)
        if p.apply_bn_for_resampling:
          t = batch_norm_inference(t, w, scope + '/bn', eps)
        t = self.store(t)
      return t

    if h > target_h and wd > target_w:
      if not p.conv_after_downsample:
        feat = maybe_1x1(feat)
      sh, sw = (h - 1) // target_h + 1, (wd - 1) // target_w + 1
      feat = max_pool_same(feat, (sh + 1, sw + 1), (sh, sw))
      if p.conv_after_downsample:
        feat = maybe_1x1(feat)
    elif h <= target_h and wd <= target_w:
      feat = maybe_1x1(feat)
      if h < target_h or wd < target_w:
        feat = resize_nearest_tf1(feat, target_h, target_w)
    else:
      raise ValueError('Incompatible target feature map size: target_height: {},'
                       'target_width: {}'.format(target_h, target_w))
    return feat

  # -- efficientdet_arch.py:418-475 -------------------------------------------------------
  def fuse_features(self, nodes, weight_method, scope):
    names = [scope + '/WSM' + ('' if i == 0 else '_%d' % i) for i in range(len(nodes))]
    return fuse_features(nodes, weight_method, [self.w[n] for n in names]
                         if weight_method != 'sum' else None)

  # -- efficientdet_arch.py:478-544 -------------------------------------------------------
  def build_bifpn_layer(self, feats, feat_sizes, rep, eps):
    p = self.p
    w = self.w
    if p.fpn_config:
      weight_method = p.fpn_config.weight_method
      nodes_cfg = [dict(n) if isinstance(n, dict) else n.as_dict() for n in p.fpn_config.nodes]
    else:
      if (p.fpn_name or 'bifpn') not in ('bifpn', 'bifpn_dyn'):
        raise NotImplementedError('fpn_name %r' % p.fpn_name)
      weight_method = p.fpn_weight_method or 'fastattn'    # tf2/fpn_configs.py:27
      nodes_cfg = [{'feat_level': lvl, 'inputs_offsets': offs}
                   for lvl, offs in structure_oracle.bifpn_nodes(p.min_level, p.max_level)]
    feats = list(feats)
    for i, fnode in enumerate(nodes_cfg):
      scope = 'fpn_cells/cell_%d/fnode%d' % (rep, i)
      th, tw = feat_sizes[fnode['feat_level']]
      nodes = []
      for idx, off in enumerate(fnode['inputs_offsets']):
        nodes.append(self.resample_feature_map(
            feats[off], '%s/resample_%d_%d_%d' % (scope, idx, off, len(feats)), th, tw, eps))
      new_node = self.fuse_features(nodes, weight_method, scope)
      op = '%s/op_after_combine%d' % (scope, len(feats))
      new_node = activation_fn(new_node, p.act_type)
      new_node = depthwise_conv2d_same(new_node, w[op + '/conv/depthwise_kernel'])
      new_node = self.store(new_node)  # engine stores the dw output before the 1x1
      new_node = conv2d_same(new_node, w[op + '/conv/pointwise_kernel']) + w[op + '/conv/bias'].view(1, -1, 1, 1)
      new_node = self.store(batch_norm_inference(new_node, w, op + '/bn', eps))
      feats.append(new_node)
    out = {}
    for l in range(p.min_level, p.max_level + 1):
      for i, fnode in enumerate(reversed(nodes_cfg)):
        if fnode['feat_level'] == l:
          out[l] = feats[-1 - i]
          break
    return out

  # -- efficientdet_arch.py:352-415 -------------------------------------------------------
  def build_feature_network(self, features, eps):
    p = self.p
    feat_sizes = structure_oracle.feature_sizes(p.image_size, p.max_level)
    feats = []
    if p.min_level not in features:
      raise ValueError('features.keys ({}) should include min_level ({})'.format(
          features.keys(), p.min_level))
    for level in range(p.min_level, p.max_level + 1):
      if level in features:
        feats.append(features[level])
      else:
        h, wd = feats[-1].shape[2], feats[-1].shape[3]
        feats.append(self.store(self.resample_feature_map(
            feats[-1], 'resample_p%d' % level, (h - 1) // 2 + 1, (wd - 1) // 2 + 1, eps)))
    # utils.verify_feats_size (utils.py:552-573)
    for f, size in zip(feats, feat_sizes[p.min_level:p.max_level + 1]):
      if (f.shape[2], f.shape[3]) != tuple(size):
        raise ValueError('feats has shape {} but it should be {}'.format(
            (f.shape[2], f.shape[3]), size))
    new_feats = None
    for rep in range(p.fpn_cell_repeats):
      new_feats = self.build_bifpn_layer(feats, feat_sizes, rep, eps)
      feats = [new_feats[l] for l in range(p.min_level, p.max_level + 1)]
    return new_feats

  # -- efficientdet_arch.py:136-302 -------------------------------------------------------
  def head(self, net, feat, level, eps):
    p = self.p
    w = self.w
    scope = '%s_net' % net
    x = feat
    for i in range(p.box_class_repeats):
      name = '%s/%s-%d' % (scope, net, i)
      x = self.store(depthwise_conv2d_same(x, w[name + '/depthwise_kernel']))
      x = conv2d_same(x, w[name + '/pointwise_kernel']) + w[name + '/bias'].view(1, -1, 1, 1)
      x = batch_norm_inference(x, w, '%s/%s-%d-bn-%d' % (scope, net, i, level), eps)
      x = self.store(activation_fn(x, p.act_type))
    name = '%s/%s-predict' % (scope, net)
    x = self.store(depthwise_conv2d_same(x, w[name + '/depthwise_kernel']))
    x = conv2d_same(x, w[name + '/pointwise_kernel']) + w[name + '/bias'].view(1, -1, 1, 1)
    return self.store(x)

  # -- efficientdet_arch.py:547-577 -------------------------------------------------------
  def __call__(self, images_nhwc):
    """images_nhwc: [N,H,W,3] array/tensor. Returns (cls_outputs, box_outputs) dicts of
    NHWC torch tensors keyed by level, like the reference."""
    x = torch.as_tensor(np.asarray(images_nhwc)).to(self.dtype).permute(0, 3, 1, 2)
    x = self.store(x)
    feats, eps = self.backbone(x)
    features = {0: x}
    features.update(feats)
    fpn = self.build_feature_network(features, eps)
    for l, t in fpn.items():
      self.endpoints['fpn_%d' % l] = t
    cls_out, box_out = {}, {}
    for l in range(self.p.min_level, self.p.max_level + 1):
      cls_out[l] = self.head('class', fpn[l], l, eps).permute(0, 2, 3, 1).contiguous()
      box_out[l] = self.head('box', fpn[l], l, eps).permute(0, 2, 3, 1).contiguous()
    return cls_out, box_out


def fuse_features(nodes, weight_method, edge_vars=None):
  """efficientdet_arch.py:418-475 on torch tensors (any layout; channel_* methods expect
  the channel axis last, as in the reference's NHWC)."""
  if weight_method == 'attn':
    nw = torch.softmax(torch.stack([v.reshape(()) for v in edge_vars]), dim=0)
    return (torch.stack(nodes, dim=-1) * nw).sum(-1)
  if weight_method == 'fastattn':
    ew = [F.relu(v) for v in edge_vars]
    ws = sum(ew)
    return sum(nodes[i] * ew[i] / (ws + 0.0001) for i in range(len(nodes)))
  if weight_method == 'channel_attn':
    nw = torch.softmax(torch.stack(edge_vars, dim=-1), dim=-1)
    return (torch.stack(nodes, dim=-1) * nw).sum(-1)
  if weight_method == 'channel_fastattn':
    ew = [F.relu(v) for v in edge_vars]
    ws = sum(ew)
    return sum(nodes[i] * ew[i] / (ws + 0.0001) for i in range(len(nodes)))
  if weight_method == 'sum':
    return sum(nodes)
  raise ValueError('unknown weight_method {}'.format(weight_method))


def fp16_store(t):
  """Models a kernel writing its result to HBM as IEEE fp16 (round-to-nearest-even)."""
  return t.to(torch.float16).to(t.dtype)
