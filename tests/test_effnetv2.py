"""EfficientNet V1 / V2 backbone (BASELINE config 3): registry pins on the
CPU, network parity against the oracle on the GPU."""
import numpy as np
import pytest
import torch

from automl_b200.efficientnetv2 import effnetv2_configs
from automl_b200.efficientnetv2 import effnetv2_model

# effnetv2_model_test.py:25-48 (Keras count_params: BN moving statistics and the Dense top included)
PARAM_PINS = {
    'efficientnet-b0': 5330564, 'efficientnet-b1': 7856232, 'efficientnet-b2': 9177562,
    'efficientnet-b3': 12314268, 'efficientnet-b4': 19466816, 'efficientnet-b5': 30562520,
    'efficientnet-b6': 43265136, 'efficientnetv2-b0': 7200312, 'efficientnetv2-b1': 8212124,
    'efficientnetv2-b2': 10178374, 'efficientnetv2-b3': 14467622, 'efficientnetv2-s': 21612360,
    'efficientnetv2-m': 54431388, 'efficientnetv2-l': 119027848, 'efficientnetv2-xl': 208896832,
}


@pytest.mark.parametrize('name', sorted(PARAM_PINS))
def test_param_counts_match_reference_pins(name):
  arch = effnetv2_model.EffNetV2Arch(name)
  assert effnetv2_model.count_params(arch) == PARAM_PINS[name]


def test_v2_s_structure():
  """V2-S stages, fused blocks without SE, single-conv expand_ratio 1 blocks."""
  a = effnetv2_model.EffNetV2Arch('efficientnetv2-s')
  assert a.stem_filters == 24 and a.head_filters == 1280 and len(a.blocks) == 40
  assert [b.conv_type for b in a.blocks[:10]] == [1] * 10 and all(b.conv_type == 0 for b in a.blocks[10:])
  assert all(b.se_filters == 0 for b in a.blocks[:10])
  assert a.blocks[10].se_filters == 16 and a.blocks[10].mid_filters == 256 and a.blocks[10].strides == 2
  assert a.blocks[0].expand_ratio == 1 and a.blocks[0].has_skip and not a.blocks[2].has_skip
  assert [a.blocks[i].output_filters for i in a.reductions] == [24, 48, 64, 160, 256]
  cfg = effnetv2_configs.get_model_config('efficientnetv2-s')
  assert cfg.eval.isize == 384 and cfg.model.act_fn == 'silu' and cfg.model.bn_epsilon == 1e-3
  with pytest.raises(ValueError):
    effnetv2_configs.get_model_config('resnet50')


def test_block_decoder_grammar():
  b = effnetv2_configs.BlockDecoder().decode(['r4_k3_s2_e4_i24_o48_c1', 'r6_k3_s2_e4_i64_o128_se0.25'])
  assert (b[0].num_repeat, b[0].kernel_size, b[0].strides, b[0].expand_ratio, b[0].input_filters,
          b[0].output_filters, b[0].conv_type, b[0].se_ratio) == (4, 3, 2, 4, 24, 48, 1, None)
  assert b[1].conv_type == 0 and b[1].se_ratio == 0.25


def test_round_filters_has_no_ninety_percent_rule():
  """effnetv2_model.py:84-95 differs from the V1 builder: no `< 0.9 * filters` bump."""
  m = effnetv2_configs.get_model_config('efficientnetv2-b2').model     # width 1.1
  assert effnetv2_model.round_filters(32, m) == 32      # 35.2 -> 32 (the V1 rule would give 40)
  assert effnetv2_model.round_filters(112, m) == 120
  assert effnetv2_model.round_repeats(5, 1.2) == 6


def rel_l2(a, b):
  a, b = a.double().flatten(), b.double().flatten()
  return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize('name,size,batch,tol', [('efficientnetv2-s', 96, 2, 2e-3),
                                                 ('efficientnetv2-b0', (64, 80), 2, 1e-3),
                                                 ('efficientnet-b0', 64, 1, 1e-3)])
def test_backbone_parity_vs_oracle(name, size, batch, tol):
  """Every block output, the reduction endpoints and the 1x1 head feature map against the fp32
  oracle on the same seeded weights and inputs: relative L2 <= 1e-3 per tensor.  The 40-block
  V2-S accumulates the fp16 rounding of 40 residual-stream tensors (the oracle's own fp16-storage
  model gives 1.0e-3 at the last block with these random weights), so its deepest tensors are
  held to 2e-3 and the first two stages to 1e-3 (open item)."""
  from oracle import effnetv2_oracle
  arch = effnetv2_model.EffNetV2Arch(name)
  w = effnetv2_model.synthetic_weights(arch, 11)
  model = effnetv2_model.get_model(name, weights=w, batch_size=batch, image_size=size)
  h, wd = model.image_size
  x = np.random.default_rng(3).uniform(-1, 1, size=(batch, h, wd, 3)).astype(np.float32)
  outs = model(torch.from_numpy(x), with_endpoints=True)
  torch.cuda.synchronize()
  ref = effnetv2_oracle.EffNetV2Oracle(arch, w, torch.float32)(x)
  worst = 0.0
  for key, t in model.endpoints.items():
    err = rel_l2(t.float().cpu().permute(0, 3, 1, 2), ref[key])
    worst = max(worst, err)
    shallow = key in ('stem', 'reduction_1', 'reduction_2') or key in ['block_%d' % i for i in range(6)]
    assert err < (1e-3 if shallow else tol), (key, err)
  assert len(outs) == 6 and outs[0].shape[-1] == arch.head_filters
  for i in range(1, 6):
    assert outs[i] is model.endpoints['reduction_%d' % i]
  again = model(torch.from_numpy(x)).clone()
  assert torch.equal(again, outs[0])                    # graph replay is deterministic
  print('%s worst rel-L2 %.2e' % (name, worst))


@pytest.mark.gpu
def test_serve_stream_equals_synchronous_calls():
  """serve_stream keeps the H2D copy of batch i+1 and the D2H copy of result i-1 under the network
  of batch i: five different batches must come back in order, equal to the synchronous calls."""
  name = 'efficientnetv2-b0'
  arch = effnetv2_model.EffNetV2Arch(name)
  w = effnetv2_model.synthetic_weights(arch, 5)
  model = effnetv2_model.get_model(name, weights=w, batch_size=2, image_size=64)
  rng = np.random.default_rng(9)
  batches = [torch.from_numpy(rng.uniform(-1, 1, size=(2, 64, 64, 3)).astype(np.float32)).pin_memory()
             for _ in range(5)]
  want = []
  for b in batches:
    want.append(model(b).cpu().clone())
  got = [r.clone() for r in model.serve_stream(batches)]     # clone: the pinned buffers are reused
  assert len(got) == 5
  for g, e in zip(got, want):
    assert torch.equal(g, e)
  assert [tuple(r.shape) for r in model.serve_stream(iter(batches[:1]))] == [tuple(want[0].shape)]
