"""Runs a few eager (un-captured) reverse steps of the bench workload so that
ncu sees every launch individually.  `python scripts/profile_step.py --batch
512 --steps 3`."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--batch', type=int, default=512)
  ap.add_argument('--steps', type=int, default=3)
  args = ap.parse_args()
  dev = torch.device('cuda', 0)
  model, cfg = bench.build_model(bench.WORKLOADS['c3'], args.batch, dev, 1)
  B, L = args.batch, bench.L
  cond = torch.full((B,), 1, dtype=torch.long, device=dev)
  xt0 = model._sample_prior(B, L)
  sig, mct, mcs = model._step_scalars(1000, 1e-5)
  loop = model._get_loop(B, L, 'cfg', cfg.guidance)
  model.use_cuda_graph = False
  # tables + labels without capturing a graph
  loop.tables = tuple(t.to(dev, torch.float32) for t in (sig, mct, mcs))
  loop.steps = sig.numel()
  loop.labels.copy_(torch.tensor([1, 2]))
  loop.xt.copy_(xt0)
  loop.step.zero_()
  loop.seed_dev.fill_(1)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  for i in range(args.steps):
    if i == args.steps - 1:
      e0.record()
    loop._body()
  e1.record()
  torch.cuda.synchronize()
  print(f'last eager step: {e0.elapsed_time(e1):.3f} ms; ids range '
        f'{int(loop.xt.min())}..{int(loop.xt.max())}')


if __name__ == '__main__':
  main()
