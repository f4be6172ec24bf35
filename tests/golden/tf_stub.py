"""A recording stand-in for `tensorflow` / `absl` (TEST INFRASTRUCTURE, used only by
tests/golden/make_structure_golden.py in the dev container).

TensorFlow cannot be installed offline, but the reference's model classes only need
`tf.keras.layers.*` as base classes / layer constructors while they BUILD the network.  With
every attribute of the stub modules resolving to a permissive class whose constructor logs
(class name, positional args, keyword args), the reference's own constructors run unmodified:

  /root/reference/efficientdet/backbone/efficientnet_model.py  Model.__init__ -> _build()
  /root/reference/efficientdet/tf2/efficientdet_keras.py       EfficientDetNet.__init__

so the REAL reference resolves its block strings, width/depth rounding, SE widths, layer names,
BiFPN node lists and head layout, and we read the result back (LOG + object attributes).
No arithmetic runs: this pins structure, not numerics.
"""
import sys
import types

LOG = []


class _Meta(type):

  def __getattr__(cls, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return _mk(cls.__name__ + '.' + name)


def _clean(v):
  if isinstance(v, (bool, int, str)) or v is None:
    return v
  if isinstance(v, float):
    return float(v)
  if isinstance(v, (list, tuple)):
    return [_clean(x) for x in v]
  try:
    import numpy as np
    if isinstance(v, np.generic):
      return v.item()
  except ImportError:
    pass
  return '<%s>' % type(v).__name__


def _init(self, *a, **k):
  LOG.append((type(self).__name__, [_clean(x) for x in a], {kk: _clean(vv) for kk, vv in k.items()}))


def _getattr(self, name):
  if name.startswith('__'):
    raise AttributeError(name)
  return _mk(name)


def _call(self, *a, **k):
  return _mk('result')()


_CACHE = {}


def _mk(name):
  if name not in _CACHE:
    _CACHE[name] = _Meta(name, (object,), {
        '__init__': _init, '__getattr__': _getattr, '__call__': _call,
        '__iter__': lambda s: iter(())})
  return _CACHE[name]


def _module(name):
  m = types.ModuleType(name)
  m.__getattr__ = lambda n: _mk(name.split('.')[-1] + '.' + n)
  m.__path__ = []
  return m


STUBBED = ['tensorflow', 'tensorflow.compat', 'tensorflow.compat.v1', 'tensorflow.compat.v2',
           'tensorflow.python', 'tensorflow.python.tpu', 'tensorflow_addons', 'tensorflow_hub',
           'tensorflow_model_optimization', 'absl', 'absl.logging', 'absl.flags']


def install():
  for n in STUBBED:
    if n not in sys.modules:
      sys.modules[n] = _module(n)
