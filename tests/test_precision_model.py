"""The format model used by the GPU parity tests of the deep / ill-conditioned networks
(tests/precision_model.py), checked on the CPU: the BN fold it applies is exact, and on a
well-conditioned network it predicts an error inside the 1e-3 bar."""
import numpy as np
import torch

from automl_b200 import arch
from automl_b200 import hparams_config
from automl_b200 import weights
from oracle import efficientdet_oracle as eo
import precision_model as pm


def _setup(name, size, seed=0, **over):
  c = hparams_config.get_efficientdet_config(name)
  c.override(dict(image_size=size, **over))
  a = arch.DetArch(c)
  w = weights.synthetic_weights(a, seed)
  x = np.random.default_rng(seed + 1).uniform(-2.0, 2.0, size=(1, size, size, 3)).astype(np.float32)
  return c, a, w, x


def test_bn_fold_of_the_model_is_exact():
  c, a, w, x = _setup('efficientdet-d0', 64)
  cls_a, box_a = eo.Oracle(c, w, torch.float64)(x)
  cls_b, box_b = eo.Oracle(c, pm.device_weights(a, w, round_gemm_weights=False), torch.float64)(x)
  for l in a.levels:
    assert pm.DeviceModel.rel_l2(cls_b[l], cls_a[l]) < 2e-6    # float32 storage of the folded values
    assert pm.DeviceModel.rel_l2(box_b[l], box_a[l]) < 2e-6


def test_format_model_on_d0_is_inside_the_bar():
  c, a, w, x = _setup('efficientdet-d0', 128)
  m = pm.DeviceModel(c, a, w, x)
  worst = max([m.endpoint_error(b.name) for b in a.blocks] +
              [m.cls_error(l) for l in a.levels] + [m.box_error(l) for l in a.levels])
  assert 1e-4 < worst < 1e-3


def test_format_model_explains_the_ill_conditioned_sum_draw():
  """Seed 3 of the 'sum' fusion test: the format alone costs > 1e-3 on a box output, seeds 0-2 do not."""
  errs = {}
  for seed in (0, 3):
    c, a, w, x = _setup('efficientdet-d1', 128, seed=seed, fpn_weight_method='sum')
    m = pm.DeviceModel(c, a, w, x)
    errs[seed] = max(m.box_error(l) for l in a.levels)
  assert errs[0] < 8e-4 < 1e-3 < errs[3]


def test_effnetv2_fold_is_exact_and_model_is_small_on_b0():
  from automl_b200.efficientnetv2 import effnetv2_model
  from oracle import effnetv2_oracle
  a = effnetv2_model.EffNetV2Arch('efficientnetv2-b0')
  w = effnetv2_model.synthetic_weights(a, 11)
  x = np.random.default_rng(3).uniform(-1, 1, size=(1, 64, 64, 3)).astype(np.float32)
  errs, ref = pm.effnetv2_format_errors(a, w, x)
  assert set(errs) == set(ref) and 1e-4 < max(errs.values()) < 1.5e-3
  # in float64 the fold is exact, so rounding the folded GEMM weights to fp16 is the only difference
  # left between the two networks: small but non-zero
  folded = pm.effnetv2_device_weights(a, w)
  ref64 = effnetv2_oracle.EffNetV2Oracle(a, w, torch.float64)(x)
  mod64 = effnetv2_oracle.EffNetV2Oracle(a, folded, torch.float64)(x)
  e = pm.DeviceModel.rel_l2(mod64['head_1x1'], ref64['head_1x1'])
  assert 1e-5 < e < 1e-3
