"""Generates tests/golden/structure.json.gz: the network STRUCTURE as resolved by the REAL
reference constructors, run in this container under the recording TensorFlow stand-in
(tests/golden/tf_stub.py; TensorFlow itself is not installable offline).

For every registered detector it instantiates /root/reference/efficientdet/tf2/
efficientdet_keras.py::EfficientDetNet(model_name) -- which builds the real
backbone/efficientnet_model.py::Model -- and records

  blocks : per MBConv block, the reference's own resolved block args (kernel, stride, input /
           output filters, expand ratio, se ratio, id_skip) and whether it built an SE module;
  layers : every Keras layer the constructors created, in creation order:
           [kind, filters, kernel, stride, use_bias, name]   (Conv2D, DepthwiseConv2D,
           SeparableConv2D, BatchNormalization);
  fnodes : per BiFPN cell and node: feat_level, inputs_offsets, weight_method, filters;
  resample: the P6.. ResampleFeatureMap layers (feat_level, channels, apply_bn);

and, under the key '__anchors__', the output of the REAL tf2/anchors.py::Anchors (numpy code
whose only TensorFlow call is the final convert_to_tensor, mapped to np.asarray(float32)) for
several (levels, scales, image size) cases: shape, sha256 of the float32 bytes, every 997th row;
and under '__feat_sizes__' the real utils.get_feat_sizes.

tests/test_structure_pins.py holds the oracle's own walk (oracle/structure_oracle.py) and the
product's DetArch to exactly this.  Run from the repo root:
  python tests/golden/make_structure_golden.py
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/efficientdet'

MODELS = ['efficientdet-d0', 'efficientdet-d1', 'efficientdet-d2', 'efficientdet-d3',
          'efficientdet-d4', 'efficientdet-d5', 'efficientdet-d6', 'efficientdet-d7',
          'efficientdet-d7x', 'efficientdet-lite0', 'efficientdet-lite1', 'efficientdet-lite2',
          'efficientdet-lite3', 'efficientdet-lite3x', 'efficientdet-lite4']

KINDS = {'Conv2D': 'conv', 'DepthwiseConv2D': 'dw', 'SeparableConv2D': 'sep',
         'BatchNormalization': 'bn'}


def _one(v):
  if isinstance(v, (list, tuple)):
    assert len(set(v)) == 1, v
    return v[0]
  return v


def canonical_layers(log):
  out = []
  for cls, args, kw in log:
    kind = KINDS.get(cls.split('.')[-1])
    if kind is None:
      continue
    if kind == 'bn':
      out.append(['bn', None, None, None, None, kw.get('name')])
      continue
    filters = kw.get('filters', args[0] if args else None)
    out.append([kind, filters, _one(kw.get('kernel_size')), _one(kw.get('strides', 1)),
                bool(kw.get('use_bias', True)), kw.get('name')])
  return out


ANCHOR_CASES = [
    # min_level, max_level, num_scales, aspect_ratios, anchor_scale, image_size
    (3, 7, 3, [1.0, 2.0, 0.5], 4.0, 512), (3, 7, 3, [1.0, 2.0, 0.5], 4.0, 640),
    (3, 7, 3, [1.0, 2.0, 0.5], 4.0, 1024), (3, 8, 3, [1.0, 2.0, 0.5], 4.0, 1536),
    (3, 7, 3, [1.0, 2.0, 0.5], 5.0, 1536), (3, 7, 3, [1.0, 2.0, 0.5], 3.0, 320),
    (3, 7, 3, [1.0, 2.0, 0.5], 4.0, (511, 513)), (3, 7, 3, [1.0, 2.0, 0.5], 4.0, '1280x640'),
    (3, 7, 2, [[1.0, 1.0], [1.4, 0.7]], [4.0, 4.0, 3.0, 3.0, 2.0], 256), (1, 1, 1, [1.0], 1.0, 8),
]


def anchors_golden():
  import hashlib
  import numpy as np
  import tensorflow as tf  # the stand-in
  tf.convert_to_tensor = lambda x, dtype=None: np.asarray(x, np.float32)
  from tf2 import anchors  # pylint: disable=g-import-not-at-top
  rows = []
  for case in ANCHOR_CASES:
    a = anchors.Anchors(*case)
    b = np.ascontiguousarray(a.boxes, np.float32)
    rows.append({'case': list(case), 'shape': list(b.shape),
                 'sha256': hashlib.sha256(b.tobytes()).hexdigest(),
                 'sample': b[::997].tolist(), 'per_location': a.get_anchors_per_location()})
  return rows


def feat_sizes_golden():
  import utils  # the reference's utils.py (pure Python helpers run as they are)
  rows = []
  for size, max_level in [(512, 7), (640, 7), (1536, 8), ((511, 513), 7), ('1280x640', 7), (8, 1)]:
    fs = utils.get_feat_sizes(size, max_level)
    rows.append({'image_size': list(size) if isinstance(size, tuple) else size,
                 'max_level': max_level, 'sizes': [[f['height'], f['width']] for f in fs]})
  return rows


def main():
  sys.path.insert(0, HERE)
  import tf_stub
  tf_stub.install()
  sys.path.insert(0, REF)
  from tf2 import efficientdet_keras as ek  # pylint: disable=g-import-not-at-top

  out = {}
  for name in MODELS:
    del tf_stub.LOG[:]
    m = ek.EfficientDetNet(model_name=name)
    log = list(tf_stub.LOG)
    blocks = []
    for b in m.backbone._blocks:  # pylint: disable=protected-access
      a = b._block_args  # pylint: disable=protected-access
      blocks.append({
          'kernel_size': a.kernel_size, 'stride': _one(a.strides),
          'input_filters': a.input_filters, 'output_filters': a.output_filters,
          'expand_ratio': a.expand_ratio, 'se_ratio': a.se_ratio, 'id_skip': bool(a.id_skip),
          'has_se': bool(b._has_se)})  # pylint: disable=protected-access
    fnodes = []
    for cell in m.fpn_cells.cells:
      fnodes.append([[fn.feat_level, list(fn.inputs_offsets), fn.weight_method,
                      fn.fpn_num_filters] for fn in cell.fnodes])
    resample = [[r.feat_level, r.target_num_channels, bool(r.apply_bn),
                 bool(r.conv_after_downsample)] for r in m.resample_layers]
    names = [kw.get('name') for cls, _, kw in log if cls in ('ResampleFeatureMap', 'FPNCell',
                                                             'FNode', 'ClassNet', 'BoxNet',
                                                             'MBConvBlock', 'MBConvBlockWithoutDepthwise')]
    out[name] = {'blocks': blocks, 'layers': canonical_layers(log), 'fnodes': fnodes,
                 'resample': resample, 'scopes': names}
    print(name, len(blocks), 'blocks', len(out[name]['layers']), 'layers')
  out['__anchors__'] = anchors_golden()
  out['__feat_sizes__'] = feat_sizes_golden()
  path = os.path.join(os.environ.get('STRUCTURE_GOLDEN_OUT', HERE), 'structure.json.gz')
  with gzip.GzipFile(path, 'wb', mtime=0) as f:
    f.write(json.dumps(out, sort_keys=True).encode())
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
