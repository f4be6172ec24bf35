"""Dumps a TensorFlow checkpoint of the reference (TF1 graph checkpoints, e.g. the
efficientdet-d0.tar.gz files linked from efficientdet/README.md:67-75) to the .npz that
automl_b200.inference.ServingDriver(ckpt_path=...) loads.

Needs an environment WITH TensorFlow (this repo's containers have none):

  python scripts/export_tf_checkpoint_to_npz.py /path/to/efficientdet-d0 efficientdet-d0.npz

Every variable is written under its checkpoint name, EMA shadows
(`.../ExponentialMovingAverage`) included; `inference.load_weights` prefers the shadows exactly like
the reference's restore_ckpt (inference.py:193-230).  Optimizer slots are skipped.
"""
import sys

import numpy as np


def main(ckpt, out):
  import tensorflow as tf  # pylint: disable=g-import-not-at-top
  if tf.io.gfile.isdir(ckpt):
    ckpt = tf.train.latest_checkpoint(ckpt)
  reader = tf.train.load_checkpoint(ckpt)
  arrays = {}
  for name in reader.get_variable_to_shape_map():
    if any(s in name for s in ('/Momentum', '/RMSProp', '/Adam', 'global_step', '_CHECKPOINTABLE')):
      continue
    arrays[name] = reader.get_tensor(name)
  np.savez(out, **arrays)
  print('wrote %d variables to %s' % (len(arrays), out))


if __name__ == '__main__':
  main(sys.argv[1], sys.argv[2])
