"""Whole-batch use_cache rule at small batch: how often does a step's forward get skipped, and do
the ids equal a run with sampling.use_cache=False whenever nothing was skipped?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
wl = bench.WORKLOADS['c3']
for B in (int(a) for a in (sys.argv[1:] or ['64'])):
  run = bench.GraphRun(wl, B, dev, 1)
  lp = run.loop
  print('B', B, 'cached mode', lp.cached)
  skipped, changes = 0, []
  prev = lp.xt.clone()
  for k in range(300):
    lp.replay(1)
    torch.cuda.synchronize()
    nl = int(lp.n_live.item()) if lp.cached else -1
    ch = int((lp.xt != prev).sum())
    prev = lp.xt.clone()
    if nl == 0:
      skipped += 1
    if k < 12 or k % 50 == 0:
      print(k, 'n_live', nl, 'tokens changed', ch, 'flags', int(lp.changed.sum()) if lp.cached else -1)
    changes.append(ch)
  print('skipped forwards', skipped, 'of 300; mean changed tokens/step', sum(changes) / len(changes))
