"""Kernel-time breakdown of one Amazon-polarity-shaped (C5: DiT-small denoiser, L=128,
V=30522, small-classifier) first-order D-CBG step (torch profiler / CUPTI).
`python scripts/profile_cbg_amazon.py [B]`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import torch

import bench
import ddg_b200

OVER = ['parameterization=d3pm', 'diffusion=uniform', 'time_conditioning=True', 'T=0',
        'model=small', 'model.length=128', 'classifier_model=small-classifier',
        'classifier_backbone=dit', 'guidance=cbg', 'guidance.gamma=2.0', 'guidance.condition=1',
        'guidance.use_approx=True', 'data.num_classes=2', 'training.guidance=null']


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
  steps = 4
  dev = torch.device('cuda', 0)
  cfg = ddg_b200.compose(OVER + [f'sampling.batch_size={B}', f'sampling.steps={steps}'])
  tok = bench.BertTok()
  model = ddg_b200.Diffusion(cfg, tok, device=dev).load_state_dict(
    ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=1))
  model.set_classifier(ddg_b200.Classifier(cfg, tok, device=dev).load_state_dict(
    ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=2, classifier=True)))
  model.sample()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  model.sample()
  torch.cuda.synchronize()
  print(f'B={B}: {(time.perf_counter() - t0) / steps * 1e3:.2f} ms/step wall')
  with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    model.sample()
    torch.cuda.synchronize()
  rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
  tot = sum(e.device_time_total for e in rows)
  print(f'total device time {tot / 1e3 / steps:.2f} ms/step over {steps} steps')
  for e in rows[:22]:
    print(f'{e.device_time_total / 1e3 / steps:9.3f} ms/step  {e.count / steps:6.1f} launches/step  {e.key[:110]}')


if __name__ == '__main__':
  main()
