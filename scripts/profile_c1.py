"""Kernel-time breakdown of one text8-shaped (C1: DiT-tiny, L=256, V=35) UDLM step at a
throughput batch (torch profiler / CUPTI).  `python scripts/profile_c1.py [B]`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import config_sweep
import ddg_b200


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
  steps = 16
  dev = torch.device('cuda', 0)
  over = ['parameterization=d3pm', 'diffusion=uniform', 'time_conditioning=True', 'T=0',
          'model=tiny', 'model.length=256', 'training.guidance=null']
  cfg = ddg_b200.compose(over + [f'sampling.batch_size={B}', f'sampling.steps={steps}'])
  tok = config_sweep.Text8Tok()
  model = ddg_b200.Diffusion(cfg, tok, device=dev).load_state_dict(
    ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=1))
  model.sample()
  torch.cuda.synchronize()
  with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    model.sample()
    torch.cuda.synchronize()
  rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
  tot = sum(e.device_time_total for e in rows)
  print(f'total device time {tot / 1e3 / steps:.2f} ms/step over {steps} steps')
  for e in rows[:20]:
    print(f'{e.device_time_total / 1e3 / steps:9.3f} ms/step  {e.count / steps:6.1f} launches/step  {e.key[:110]}')


if __name__ == '__main__':
  main()
