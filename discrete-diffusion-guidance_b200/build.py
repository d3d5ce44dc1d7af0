"""Builds csrc/*.cu into libddg_b200.so (in-tree, next to this file) with nvcc
for sm_100a.  nvcc cross-compiles without a GPU; the .so travels with the repo
snapshot to the GPU box."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libddg_b200.so')
STAMP = os.path.join(HERE, 'csrc', '.build_stamp')
SOURCES = ['capi.cu', 'sample_step.cu', 'sample_fused.cu', 'sample_group.cu', 'sample_aux.cu', 'dit_ops.cu', 'dit_bwd.cu', 'attention_bwd.cu', 'attention_bwd_tc.cu',
           'gemm_tcgen05.cu', 'attention.cu']
NVCC_FLAGS = [
  '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
  '-std=c++17', '--threads', '8', '-Xcompiler', '-fPIC', '-Xcompiler',
  '-fvisibility=hidden', '--shared']


def _digest() -> str:
  h = hashlib.sha256()
  files = sorted(os.listdir(CSRC))
  paths = [os.path.join(CSRC, f) for f in files if f.endswith(('.cu', '.cuh'))]
  paths.append(os.path.join(HERE, '..', 'include', 'ddg_b200.h'))
  for p in paths:
    with open(p, 'rb') as f:
      h.update(f.read())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
  """Compiles the library if sources changed. Returns the .so path."""
  dig = _digest()
  if (not force and os.path.exists(LIB) and os.path.exists(STAMP)
      and open(STAMP).read().strip() == dig):
    return LIB
  nvcc = os.environ.get('NVCC', 'nvcc')
  cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + [
    '-o', LIB]
  if verbose:
    print(' '.join(cmd), file=sys.stderr)
  subprocess.run(cmd, check=True, cwd=HERE)
  with open(STAMP, 'w') as f:
    f.write(dig)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose=True))
