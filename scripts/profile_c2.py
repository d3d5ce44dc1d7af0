"""Kernel-time breakdown of the MDLM loop (BASELINE.json configs[1]: per-sequence logits cache) over
a window of graph-replayed steps in the middle of a 1000-step run.
`python scripts/profile_c2.py [first_step] [n_steps] [B]`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
  first = int(sys.argv[1]) if len(sys.argv) > 1 else 400
  n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
  wl = bench.WORKLOADS['c2']
  B = int(sys.argv[3]) if len(sys.argv) > 3 else wl.batch
  dev = torch.device('cuda', 0)
  run = bench.GraphRun(wl, B, dev, 1)
  run.replay(first)
  torch.cuda.synchronize()
  live = int(run.loop.n_live.item()) if run.loop.cached else -1
  masked = int((run.loop.xt == run.model.mask_index).sum())
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  run.replay(n)
  e1.record()
  torch.cuda.synchronize()
  print(f'c2 B={B}: steps {first}..{first + n}: {e0.elapsed_time(e1) / n:.3f} ms/step; live sequences at '
        f'step {first}: {live}, masked tokens {masked} of {B * run.Lw}')
  with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    run.replay(n)
    torch.cuda.synchronize()
  rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
  tot = sum(e.device_time_total for e in rows)
  print(f'total device time {tot / 1e3 / n:.3f} ms/step over {n} steps')
  for e in rows[:16]:
    print(f'{e.device_time_total / 1e3 / n:9.4f} ms/step  {e.count / n:6.1f} launches/step  {e.key[:110]}')


if __name__ == '__main__':
  main()
