"""CPU checks of the C-ABI boundary: the shared library loads, exports every
symbol include/ddg_b200.h declares, and refuses to compute without a B200."""
import ctypes
import os
import re

import pytest
import torch

import ddg_b200
from ddg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  src = open(os.path.join(ROOT, 'include', 'ddg_b200.h')).read()
  return sorted(set(re.findall(r'DDG_API\s+[\w\s\*]+?\b(ddg_\w+)\s*\(', src)))


def test_header_symbols_are_exported():
  names = _declared()
  assert len(names) >= 16
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for n in names:
    assert hasattr(lib, n), f'{n} declared in ddg_b200.h but not exported'
  assert set(names) == set(_lib.EXPORTS)
  assert lib.ddg_abi_version() == 6 == ddg_b200._lib.ABI_VERSION


def test_struct_layouts_match_header():
  """ctypes mirrors of the argument structs have the C sizes (checked against
  a compiled sizeof)."""
  import subprocess
  import tempfile
  prog = ('#include "ddg_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n",'
          'sizeof(ddg_sample_step_args), sizeof(ddg_gemm_args));return 0;}')
  with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, 't.c'), 'w').write(prog)
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 't.c'),
                    '-o', os.path.join(d, 't')], check=True)
    out = subprocess.run([os.path.join(d, 't')], capture_output=True, text=True, check=True)
  s1, s2 = map(int, out.stdout.split())
  assert ctypes.sizeof(_lib.SampleStepArgs) == s1
  assert ctypes.sizeof(_lib.GemmArgs) == s2


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_no_cpu_fallback():
  with pytest.raises(RuntimeError):
    _lib.require_device()
  a = _lib.SampleStepArgs(rows=1, vocab=8, seq_len=1)
  assert _lib.lib.ddg_sample_step(a, 0) == _lib.ERR_UNSUPPORTED
  g = _lib.GemmArgs(M=1, N=8, K=8)
  assert _lib.lib.ddg_gemm(g, 0) == _lib.ERR_UNSUPPORTED
  with pytest.raises(RuntimeError):
    ddg_b200.Diffusion(ddg_b200.compose([]), None)
