"""Kernel-time breakdown of one QM9-shaped D-CBG step (torch profiler / CUPTI).
`python scripts/profile_cbg.py [B] [exact]`  (default: first-order, B = 4096)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import config_sweep
import ddg_b200


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
  dev = torch.device('cuda', 0)
  exact = len(sys.argv) > 2 and sys.argv[2] == 'exact'
  over, _, steps = config_sweep.CBG_CONFIGS[
    'C4 QM9 cbg exact B=64' if exact else 'C4 QM9 cbg first-order B=4096']
  cfg = ddg_b200.compose(list(over) + [f'sampling.batch_size={B}', f'sampling.steps={steps}'])
  tok = config_sweep.Qm9Tok()
  model = ddg_b200.Diffusion(cfg, tok, device=dev).load_state_dict(
    ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=1))
  model.set_classifier(ddg_b200.Classifier(cfg, tok, device=dev).load_state_dict(
    ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=2, classifier=True)))
  model.sample()
  torch.cuda.synchronize()
  with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    model.sample()
    torch.cuda.synchronize()
  rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
  tot = sum(e.device_time_total for e in rows)
  print(f'total device time {tot / 1e3 / steps:.2f} ms/step over {steps} steps')
  for e in rows[:25]:
    print(f'{e.device_time_total / 1e3 / steps:9.3f} ms/step  {e.count / steps:6.1f} launches/step  {e.key[:110]}')


if __name__ == '__main__':
  main()
