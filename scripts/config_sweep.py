"""Steady-state step time of the other BASELINE.json configs' shapes (random-init weights,
Philox noise): a sanity sweep, not the headline bench.  `python scripts/config_sweep.py`"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import ddg_b200

CONFIGS = {
  'C2 LM1B MDLM subs null B=512': (['parameterization=subs', 'diffusion=absorbing_state',
                                    'time_conditioning=False', 'model=small', 'model.length=128',
                                    'training.guidance=null'], 512),
  'C3 LM1B UDLM cfg g=2 B=512': (bench.OVERRIDES, 512),
  'C3 at 64 / GPU': (bench.OVERRIDES, 64),
  'C5 UDLM cfg g=2 B=1024': (bench.OVERRIDES, 1024),
}


def main():
  dev = torch.device('cuda', 0)
  steps = 40
  for name, (over, B) in CONFIGS.items():
    cfg = ddg_b200.compose(list(over) + [f'sampling.batch_size={B}', f'sampling.steps={steps}'])
    sd = ddg_b200.weights.random_state_dict(cfg, bench.V, seed=1)
    model = ddg_b200.Diffusion(cfg, bench.BertTok(), device=dev).load_state_dict(sd)
    model.sample()                      # capture + warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.sample()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    masked = float((out == 103).float().mean()) if 'subs' in ' '.join(over) else 0.0
    print(f'{name:32s} {dt / steps * 1e3:8.3f} ms/step  -> {B / (1000 * dt / steps):7.2f} seq/s at 1000 steps'
          f'  (masks left {masked:.3f})', flush=True)
    del model
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
