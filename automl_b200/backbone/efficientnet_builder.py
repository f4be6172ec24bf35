"""EfficientNet / EfficientNet-lite backbone description (host side).

Turns a backbone name into the flat list of resolved MBConv blocks the CUDA
engine (and the oracle) walk. No tensors here.

Reference behaviour mirrored (file:line under /root/reference/efficientdet/backbone):
  * efficientnet_builder.py:31-46        width/depth coefficients per model
  * efficientnet_builder.py:52-79        block string mini-language (r,k,s,e,i,o,se,...)
  * efficientnet_builder.py:163-168      the seven default stages
  * efficientnet_lite_builder.py:33-79   lite variants: relu6, no SE, fixed stem/head
  * efficientnet_model.py:128-150        round_filters (divisor 8, 90% rule) / round_repeats
  * efficientnet_model.py:639-705        stage -> repeated blocks expansion
  * efficientnet_model.py:269-358        per-block layer names (conv2d[_n], tpu_batch_normalization[_n])
  * efficientnet_model.py:738-759        which block outputs are reduction_1..5
"""
import collections
import math
import re

BlockArgs = collections.namedtuple('BlockArgs', [
    'kernel_size', 'num_repeat', 'input_filters', 'output_filters',
    'expand_ratio', 'id_skip', 'strides', 'se_ratio', 'conv_type', 'fused_conv',
    'super_pixel', 'condconv'
])

GlobalParams = collections.namedtuple('GlobalParams', [
    'width_coefficient', 'depth_coefficient', 'depth_divisor', 'min_depth',
    'batch_norm_epsilon', 'use_se', 'fix_head_stem', 'local_pooling', 'act_type',
    'blocks_args'
])

# One resolved MBConv block, ready to be lowered to kernels.
BlockSpec = collections.namedtuple('BlockSpec', [
    'name',          # 'blocks_7'
    'kernel_size', 'stride', 'input_filters', 'output_filters', 'expand_ratio',
    'mid_filters',   # input_filters * expand_ratio
    'se_filters',    # 0 when the block has no SE
    'has_skip',      # identity add at the end
    'expand_name', 'expand_bn',    # None when expand_ratio == 1
    'dw_bn', 'project_name', 'project_bn',
    'reduction',     # 1..5 if this block's output is endpoint reduction_k, else 0
])

_COEFFS = {
    # name: (width_coefficient, depth_coefficient)
    'efficientnet-b0': (1.0, 1.0),
    'efficientnet-b1': (1.0, 1.1),
    'efficientnet-b2': (1.1, 1.2),
    'efficientnet-b3': (1.2, 1.4),
    'efficientnet-b4': (1.4, 1.8),
    'efficientnet-b5': (1.6, 2.2),
    'efficientnet-b6': (1.8, 2.6),
    'efficientnet-b7': (2.0, 3.1),
    'efficientnet-b8': (2.2, 3.6),
    'efficientnet-l2': (4.3, 5.3),
    'efficientnet-lite0': (1.0, 1.0),
    'efficientnet-lite1': (1.0, 1.1),
    'efficientnet-lite2': (1.1, 1.2),
    'efficientnet-lite3': (1.2, 1.4),
    'efficientnet-lite4': (1.4, 1.8),
}

_DEFAULT_BLOCKS_ARGS = [
    'r1_k3_s11_e1_i32_o16_se0.25', 'r2_k3_s22_e6_i16_o24_se0.25',
    'r2_k5_s22_e6_i24_o40_se0.25', 'r3_k3_s22_e6_i40_o80_se0.25',
    'r3_k5_s11_e6_i80_o112_se0.25', 'r4_k5_s22_e6_i112_o192_se0.25',
    'r1_k3_s11_e6_i192_o320_se0.25',
]


def efficientnet_params(model_name):
  """(width_coefficient, depth_coefficient); KeyError for unknown names."""
  return _COEFFS[model_name]


class BlockDecoder(object):
  """String <-> BlockArgs, e.g. 'r2_k5_s22_e6_i24_o40_se0.25'."""

  def _decode_block_string(self, block_string):
    assert isinstance(block_string, str)
    options = {}
    for op in block_string.split('_'):
      m = re.match(r'([a-z]+)(\d.*)$', op)
      if m:
        options[m.group(1)] = m.group(2)
    if 's' not in options or len(options['s']) != 2:
      raise ValueError('Strides options should be a pair of integers.')
    return BlockArgs(
        kernel_size=int(options['k']),
        num_repeat=int(options['r']),
        input_filters=int(options['i']),
        output_filters=int(options['o']),
        expand_ratio=int(options['e']),
        id_skip=('noskip' not in block_string),
        se_ratio=float(options['se']) if 'se' in options else None,
        strides=[int(options['s'][0]), int(options['s'][1])],
        conv_type=int(options['c']) if 'c' in options else 0,
        fused_conv=int(options['f']) if 'f' in options else 0,
        super_pixel=int(options['p']) if 'p' in options else 0,
        condconv=('cc' in block_string))

  def _encode_block_string(self, block):
    parts = [
        'r%d' % block.num_repeat,
        'k%d' % block.kernel_size,
        's%d%d' % (block.strides[0], block.strides[1]),
        'e%s' % block.expand_ratio,
        'i%d' % block.input_filters,
        'o%d' % block.output_filters,
        'c%d' % block.conv_type,
        'f%d' % block.fused_conv,
        'p%d' % block.super_pixel,
    ]
    if block.se_ratio is not None and 0 < block.se_ratio <= 1:
      parts.append('se%s' % block.se_ratio)
    if block.id_skip is False:
      parts.append('noskip')
    if block.condconv:
      parts.append('cc')
    return '_'.join(parts)

  def decode(self, string_list):
    assert isinstance(string_list, list)
    return [self._decode_block_string(s) for s in string_list]

  def encode(self, blocks_args):
    return [self._encode_block_string(b) for b in blocks_args]


def round_filters(filters, global_params, skip=False):
  """Scale by the width multiplier, round to the divisor, never lose >10%."""
  multiplier = global_params.width_coefficient
  divisor = global_params.depth_divisor
  if skip or not multiplier:
    return filters
  filters *= multiplier
  min_depth = global_params.min_depth or divisor
  new_filters = max(min_depth, int(filters + divisor / 2) // divisor * divisor)
  if new_filters < 0.9 * filters:
    new_filters += divisor
  return int(new_filters)


def round_repeats(repeats, global_params, skip=False):
  multiplier = global_params.depth_coefficient
  if skip or not multiplier:
    return repeats
  return int(math.ceil(multiplier * repeats))


def get_model_params(model_name, override_params=None):
  """(blocks_args, global_params) for efficientnet-bN / efficientnet-liteN."""
  if not model_name.startswith('efficientnet-'):
    raise ValueError('Unknown model name {}'.format(model_name))
  if model_name not in _COEFFS:
    raise NotImplementedError(
        'model name is not pre-defined: %s' % model_name)
  width, depth = _COEFFS[model_name]
  lite = model_name.startswith('efficientnet-lite')
  gp = GlobalParams(
      width_coefficient=width,
      depth_coefficient=depth,
      depth_divisor=8,
      min_depth=None,
      batch_norm_epsilon=1e-3,
      use_se=not lite,
      fix_head_stem=lite,
      local_pooling=lite,
      act_type='relu6' if lite else 'swish',
      blocks_args=_DEFAULT_BLOCKS_ARGS)
  if override_params:
    gp = gp._replace(**override_params)  # ValueError on unknown fields
  return BlockDecoder().decode(list(gp.blocks_args)), gp


def _layer_namer(prefix):
  """'conv2d', 'conv2d_1', 'conv2d_2', ... in creation order."""
  count = [0]

  def nxt():
    n = count[0]
    count[0] += 1
    return prefix if n == 0 else '%s_%d' % (prefix, n)

  return nxt


def expand_blocks(blocks_args, global_params):
  """Stage list -> (stem_filters, [BlockSpec]) with reduction endpoints marked."""
  stem_filters = round_filters(blocks_args[0].input_filters, global_params,
                               global_params.fix_head_stem)
  flat = []  # (kernel, stride, cin, cout, expand, se_ratio, id_skip)
  n_stages = len(blocks_args)
  for i, ba in enumerate(blocks_args):
    assert ba.num_repeat > 0
    if ba.super_pixel or ba.fused_conv or ba.conv_type or ba.condconv:
      raise NotImplementedError(
          'super_pixel / fused_conv / conv_type / condconv blocks are not used '
          'by any registered EfficientDet backbone')
    if ba.strides[0] != ba.strides[1]:
      raise NotImplementedError('non-square strides')
    cin = round_filters(ba.input_filters, global_params)
    cout = round_filters(ba.output_filters, global_params)
    if global_params.fix_head_stem and (i == 0 or i == n_stages - 1):
      repeats = ba.num_repeat
    else:
      repeats = round_repeats(ba.num_repeat, global_params)
    flat.append((ba.kernel_size, ba.strides[0], cin, cout, ba.expand_ratio,
                 ba.se_ratio, ba.id_skip))
    for _ in range(repeats - 1):
      flat.append((ba.kernel_size, 1, cout, cout, ba.expand_ratio, ba.se_ratio,
                   ba.id_skip))

  specs = []
  reduction_idx = 0
  for idx, (k, s, cin, cout, e, se_ratio, id_skip) in enumerate(flat):
    is_reduction = (idx == len(flat) - 1) or flat[idx + 1][1] > 1
    if is_reduction:
      reduction_idx += 1
    conv_name = _layer_namer('conv2d')
    bn_name = _layer_namer('tpu_batch_normalization')
    expand_name = expand_bn = None
    if e != 1:
      expand_name, expand_bn = conv_name(), bn_name()
    dw_bn = bn_name()
    has_se = (global_params.use_se and se_ratio is not None and
              0 < se_ratio <= 1)
    # Keras layers infer their input width from the tensor they get: with fix_head_stem (lite)
    # the stem stays at 32 filters while blocks_args[0].input_filters is width-scaled (40 for
    # lite3, 48 for lite4), so the first block's convs are built on the STEM's channel count
    # (efficientnet_model.py:512-513 vs :652-653; parameter pins efficientdet_arch_test.py:104-114).
    actual_in = stem_filters if idx == 0 else cin
    specs.append(
        BlockSpec(
            name='blocks_%d' % idx,
            kernel_size=k,
            stride=s,
            input_filters=actual_in,
            output_filters=cout,
            expand_ratio=e,
            mid_filters=actual_in if e == 1 else cin * e,
            se_filters=max(1, int(cin * se_ratio)) if has_se else 0,
            has_skip=bool(id_skip and s == 1 and cin == cout),
            expand_name=expand_name,
            expand_bn=expand_bn,
            dw_bn=dw_bn,
            project_name=conv_name(),
            project_bn=bn_name(),
            reduction=reduction_idx if is_reduction else 0))
  return stem_filters, specs


def backbone_spec(model_name, override_params=None):
  """Convenience: (global_params, stem_filters, [BlockSpec]) for a backbone name."""
  blocks_args, gp = get_model_params(model_name, override_params)
  stem_filters, specs = expand_blocks(blocks_args, gp)
  return gp, stem_filters, specs
