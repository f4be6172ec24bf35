"""Multi-scale anchor boxes (host side, numpy, built once per model).

Mirrors /root/reference/efficientdet/tf2/anchors.py:83-168 (Anchors) — float64
numpy maths cast to float32 at the end, order (level, y, x, anchor) with
anchor = octave * len(aspect_ratios) + aspect_index — and :30-58
(decode_box_outputs, here as a numpy helper used by tests; the product decode
runs in csrc/postprocess.cu).
"""
import numpy as np

from automl_b200 import utils

MAX_DETECTION_POINTS = 5000  # anchors.py:27


class Anchors(object):
  """Multi-scale anchors; `.boxes` is float32 [A_total, 4] = [ymin,xmin,ymax,xmax]."""

  def __init__(self, min_level, max_level, num_scales, aspect_ratios,
               anchor_scale, image_size):
    self.min_level = min_level
    self.max_level = max_level
    self.num_scales = num_scales
    self.aspect_ratios = aspect_ratios
    n_levels = max_level - min_level + 1
    if isinstance(anchor_scale, (list, tuple)):
      assert len(anchor_scale) == n_levels
      self.anchor_scales = anchor_scale
    else:
      self.anchor_scales = [anchor_scale] * n_levels
    self.image_size = utils.parse_image_size(image_size)
    self.feat_sizes = utils.get_feat_sizes(image_size, max_level)
    self.config = self._generate_configs()
    self.boxes = self._generate_boxes()

  def _generate_configs(self):
    fs = self.feat_sizes
    cfg = {}
    for level in range(self.min_level, self.max_level + 1):
      stride = (fs[0]['height'] / float(fs[level]['height']),
                fs[0]['width'] / float(fs[level]['width']))
      cfg[level] = [(stride, octave / float(self.num_scales), aspect,
                     self.anchor_scales[level - self.min_level])
                    for octave in range(self.num_scales)
                    for aspect in self.aspect_ratios]
    return cfg

  def _generate_boxes(self):
    per_level = []
    for level in range(self.min_level, self.max_level + 1):
      per_anchor = []
      for stride, octave_scale, aspect, anchor_scale in self.config[level]:
        base_x = anchor_scale * stride[1] * 2**octave_scale
        base_y = anchor_scale * stride[0] * 2**octave_scale
        if isinstance(aspect, list):
          aspect_x, aspect_y = aspect
        else:
          aspect_x = np.sqrt(aspect)
          aspect_y = 1.0 / aspect_x
        half_x = base_x * aspect_x / 2.0
        half_y = base_y * aspect_y / 2.0
        xs = np.arange(stride[1] / 2, self.image_size[1], stride[1])
        ys = np.arange(stride[0] / 2, self.image_size[0], stride[0])
        xv, yv = np.meshgrid(xs, ys)
        xv, yv = xv.reshape(-1), yv.reshape(-1)
        # [HW, 4] for this anchor shape
        per_anchor.append(
            np.stack([yv - half_y, xv - half_x, yv + half_y, xv + half_x], 1))
      # [HW, A, 4] -> [HW*A, 4]: location-major, anchor-minor.
      per_level.append(np.stack(per_anchor, axis=1).reshape(-1, 4))
    return np.concatenate(per_level, axis=0).astype(np.float32)

  def get_anchors_per_location(self):
    return self.num_scales * len(self.aspect_ratios)


def decode_box_outputs(pred_boxes, anchor_boxes):
  """(ty,tx,th,tw) relative to anchors -> [ymin,xmin,ymax,xmax]; float32 numpy."""
  pred_boxes = np.asarray(pred_boxes, np.float32)
  a = np.asarray(anchor_boxes, np.float32)
  ycenter_a = (a[..., 0] + a[..., 2]) / np.float32(2)
  xcenter_a = (a[..., 1] + a[..., 3]) / np.float32(2)
  ha = a[..., 2] - a[..., 0]
  wa = a[..., 3] - a[..., 1]
  ty, tx, th, tw = (pred_boxes[..., i] for i in range(4))
  w = np.exp(tw) * wa
  h = np.exp(th) * ha
  ycenter = ty * ha + ycenter_a
  xcenter = tx * wa + xcenter_a
  return np.stack([ycenter - h / np.float32(2), xcenter - w / np.float32(2),
                   ycenter + h / np.float32(2), xcenter + w / np.float32(2)],
                  axis=-1)
