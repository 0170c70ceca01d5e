"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the product never imports the oracle."""
import ast
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
  text = open(os.path.join(ROOT, 'include', 'b200rl.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(b200rl_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
  from acme_b200 import _capi
  lib = _capi.load()
  syms = header_symbols()
  assert len(syms) >= 40
  for s in syms:
    assert hasattr(lib, s), f'{s} declared in include/b200rl.h but not exported'
  assert set(syms) == set(_capi.PROTOTYPES), set(syms) ^ set(_capi.PROTOTYPES)
  assert lib.b200rl_version() == 100


def test_cubins_are_sm100a():
  import shutil
  import subprocess
  from acme_b200 import _capi
  if not shutil.which('cuobjdump'):
    pytest.skip('cuobjdump not on PATH')
  out = subprocess.run(['cuobjdump', '--list-elf', _capi._LIB_PATH], capture_output=True, text=True).stdout
  assert 'sm_100a' in out and 'sm_90' not in out


def test_fails_loudly_without_a_gpu():
  import torch
  from acme_b200 import _capi
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  with pytest.raises(_capi.B200RLError):
    _capi.require_device(0)


def test_product_does_not_import_oracle():
  pkg = os.path.join(ROOT, 'acme_b200')
  for fn in os.listdir(pkg):
    if not fn.endswith('.py'):
      continue
    treeobj = ast.parse(open(os.path.join(pkg, fn)).read())
    for node in ast.walk(treeobj):
      names = []
      if isinstance(node, ast.Import):
        names = [a.name for a in node.names]
      elif isinstance(node, ast.ImportFrom):
        names = [node.module or '']
      assert not any(n == 'oracle' or n.startswith('oracle.') for n in names), fn
