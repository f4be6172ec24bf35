"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/automl_b200.h declares (and ctypes signatures exist for all of them)."""
import ctypes
import os
import re

from automl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  with open(os.path.join(ROOT, 'include', 'automl_b200.h')) as f:
    text = f.read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(edet_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
  import __graft_entry__
  __graft_entry__.build()
  names = _declared_symbols()
  assert len(names) >= 12
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in names:
    assert hasattr(lib, name), 'missing export %s' % name
    assert name in _lib.SIGNATURES, 'no ctypes signature for %s' % name
  assert sorted(_lib.SIGNATURES) == names
  assert _lib.load().edet_version() >= 100


def test_no_product_import_of_oracle():
  """The oracle is test infrastructure: nothing under automl_b200/ may import it."""
  pkg = os.path.join(ROOT, 'automl_b200')
  for dirpath, _, files in os.walk(pkg):
    for fn in files:
      if fn.endswith('.py'):
        with open(os.path.join(dirpath, fn)) as f:
          src = f.read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
