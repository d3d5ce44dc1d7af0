"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the unmodified reference where the
GPU box can see it.

`/root/reference` exists only in the build container.  `stage()` (called by
`__graft_entry__.build()`) copies the handful of Python files the sampling path
imports into `baseline/_ref/` -- git-ignored, so nothing
of the reference enters this repo's history, but shipped with the `gpurun`
snapshot -- so that

  * `bench.py --impl reference` times the reference's OWN `Diffusion._cfg_denoise`
    on the box's host cores, and
  * `tests/test_gpu_reference.py` runs the reference itself on the B200 under its
    real cuda autocast policy and pins the CUDA path to it.

The files are byte-for-byte copies; `oracle/refshim.py` imports them where they
lie (with import-time stubs for the packages this image lacks).
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = '/root/reference'
STAGED = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['diffusion.py', 'classifier.py', 'dataloader.py', 'noise_schedule.py',
         'utils.py', 'tokenizer.py', 'eval_utils.py', 'main.py', 'LICENSE']
DIRS = ['models', 'custom_datasets', 'configs']


def stage(verbose: bool = False) -> str | None:
  """Copies the reference's sampling-path modules to baseline/_ref/.  Returns the
  staged path, or None when /root/reference is absent (GPU box: uses what the
  build container staged)."""
  if not os.path.isfile(os.path.join(SOURCE, 'diffusion.py')):
    return STAGED if os.path.isfile(os.path.join(STAGED, 'diffusion.py')) else None
  os.makedirs(STAGED, exist_ok=True)
  for f in FILES:
    src = os.path.join(SOURCE, f)
    if os.path.isfile(src):
      shutil.copyfile(src, os.path.join(STAGED, f))
  for d in DIRS:
    src = os.path.join(SOURCE, d)
    if os.path.isdir(src):
      shutil.copytree(src, os.path.join(STAGED, d), dirs_exist_ok=True,
                      ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
  if verbose:
    print(f'staged reference -> {STAGED}')
  return STAGED


if __name__ == '__main__':
  print(stage(verbose=True))
