"""ServingDriver call surface (reference inference.py:340-554) and the device pre-process."""
import numpy as np
import pytest
import torch

from oracle import postprocess_oracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('hw,size', [((96, 128), 128), ((200, 150), (128, 160)), ((64, 64), 64)])
def test_preprocess_matches_oracle(hw, size):
  from automl_b200 import ops, utils
  rng = np.random.default_rng(hw[0])
  imgs = rng.integers(0, 256, size=(2,) + hw + (3,), dtype=np.uint8)
  mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
  oh, ow = utils.parse_image_size(size)
  out = torch.empty(2, oh, ow, 3, device='cuda:0')
  scale = ops.preprocess(torch.from_numpy(imgs).cuda(), out, mean, std)
  torch.cuda.synchronize()
  for i in range(2):
    ref, ref_scale = po.image_preprocess(imgs[i], size, mean, std)
    np.testing.assert_allclose(out[i].cpu().numpy(), ref, rtol=1e-5, atol=2e-6)
    assert abs(scale - ref_scale) <= 1e-6 * ref_scale


def test_serving_driver_call_surface():
  from automl_b200 import inference
  rng = np.random.default_rng(0)
  imgs = [rng.integers(0, 256, size=(96, 128, 3), dtype=np.uint8) for _ in range(2)]
  driver = inference.ServingDriver('efficientdet-d0', '_', batch_size=2,
                                   model_params={'image_size': 128})
  assert driver.params['is_training_bn'] is False and driver.params['image_size'] == 128
  pred = driver.serve_images(imgs)            # lazily builds, like the reference
  assert sorted(driver.signitures) == ['image_arrays', 'image_files', 'prediction']
  assert pred.shape == (2, 100, 7) and pred.dtype == np.float32
  np.testing.assert_array_equal(pred[:, :, 0], [[0.0] * 100, [1.0] * 100])
  assert ((pred[:, :, 6] >= 1) & (pred[:, :, 6] <= 90)).all()      # 1-based classes
  assert (np.diff(pred[:, :, 5], axis=1) <= 0).all()               # scores sorted per image
  # boxes are clipped to the network input and scaled back to the original image
  scale = 1.0 / min(128 / 96, 128 / 128)
  assert pred[:, :, 1:5].min() >= 0 and pred[:, :, 1:5].max() <= 128 * scale + 1e-3
  # the engine input is the oracle's pre-process of the raw image
  ref, _ = po.image_preprocess(imgs[1], 128, driver.params['mean_rgb'], driver.params['stddev_rgb'])
  np.testing.assert_allclose(driver.engine.input[1].cpu().numpy(), ref, rtol=1e-5, atol=2e-6)
  again = driver.serve_images(imgs)
  np.testing.assert_array_equal(pred, again)                       # deterministic
  with pytest.raises(ValueError):
    driver.serve_images(imgs[:1])
  with pytest.raises(NotImplementedError):
    driver.export('/tmp/x')
  with pytest.raises(ValueError):
    inference.ServingDriver('resnet50', '_')
