"""Kernel-time breakdown of one D-CBG step (torch profiler / CUPTI) of a bench workload.
`python scripts/profile_cbg.py [c4|c4_exact|c5_cbg] [B]`  (default: c4 = QM9-shaped
first-order D-CBG at B = 4096; c5_cbg = Amazon-shaped: DiT-small denoiser + classifier,
L = 128, V = 30522)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
  key = sys.argv[1] if len(sys.argv) > 1 else 'c4'
  wl = bench.WORKLOADS[key]
  B = int(sys.argv[2]) if len(sys.argv) > 2 else wl.batch
  steps = 4
  dev = torch.device('cuda', 0)
  model, cfg = bench.build_model(wl, B, dev, 1, steps=steps)
  model.sample()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  model.sample()
  torch.cuda.synchronize()
  print(f'{key} B={B}: {(time.perf_counter() - t0) / steps * 1e3:.2f} ms/step wall')
  with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    model.sample()
    torch.cuda.synchronize()
  rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
  tot = sum(e.device_time_total for e in rows)
  print(f'total device time {tot / 1e3 / steps:.2f} ms/step over {steps} steps')
  for e in rows[:26]:
    print(f'{e.device_time_total / 1e3 / steps:9.3f} ms/step  {e.count / steps:6.1f} launches/step  {e.key[:120]}')


if __name__ == '__main__':
  main()
