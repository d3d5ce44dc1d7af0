"""CPU check of bench.py's reference arm: one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--steps', '1', '--warmup', '1'], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['unit'] == 'sequences/s'
  assert d['metric'].startswith('sampled sequences/sec, LM1B UDLM D-CFG')
  assert d['value'] > 0 and d['ms_per_step'] > 0 and d['n_gpus'] == 1
  assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
  assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
  assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
  assert 'workload' in d['config'] and d['vs_baseline'] is None
  # the reference arm runs the reference alone: it never loads the product package / library
  assert d['product_modules_loaded'] == [] and d['native_so_loaded'] == []


def test_reference_arm_other_ranks_stay_silent():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--gpus', '2', '--steps', '1', '--warmup', '1'], capture_output=True,
                       text=True, timeout=120, cwd=ROOT, env=env)
  assert out.returncode == 0 and out.stdout.strip() == ''
