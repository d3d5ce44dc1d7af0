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

# C4: QM9-shaped (V=40, L=32) UDLM + DiT classifier (tiny-classifier), D-CBG
QM9 = ['parameterization=d3pm', 'diffusion=uniform', 'time_conditioning=True', 'T=0',
       'model=small', 'model.length=32', 'classifier_model=tiny-classifier',
       'classifier_backbone=dit', 'guidance=cbg', 'guidance.gamma=2.0', 'guidance.condition=1',
       'data.num_classes=2', 'training.guidance=null']
CBG_CONFIGS = {
  'C4 QM9 cbg first-order B=4096': (QM9 + ['guidance.use_approx=True'], 4096, 8),
  'C4 QM9 cbg exact B=64': (QM9 + ['guidance.use_approx=False'], 64, 4),
}


class Qm9Tok:
  vocab_size, mask_token_id, bos_token_id, eos_token_id, pad_token_id = 40, 4, 2, 3, 5
  mask_token = '[MASK]'


def cbg_main(dev):
  for name, (over, B, steps) in CBG_CONFIGS.items():
    cfg = ddg_b200.compose(list(over) + [f'sampling.batch_size={B}', f'sampling.steps={steps}'])
    tok = Qm9Tok()
    sd = ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=1)
    csd = ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=2, classifier=True)
    model = ddg_b200.Diffusion(cfg, tok, device=dev).load_state_dict(sd)
    model.set_classifier(ddg_b200.Classifier(cfg, tok, device=dev).load_state_dict(csd))
    model.sample()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.sample()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert int(out.min()) >= 0 and int(out.max()) < tok.vocab_size
    print(f'{name:32s} {dt / steps * 1e3:8.3f} ms/step  -> {B / (32 * dt / steps):7.2f} seq/s at 32 steps',
          flush=True)
    del model
    torch.cuda.empty_cache()


class Text8Tok:
  vocab_size, mask_token_id, bos_token_id, eos_token_id, pad_token_id = 35, 4, 0, 1, 3
  mask_token = '[MASK]'


def c1_main(dev):
  """C1: text8 UDLM DiT-tiny, L=256, V=35, 128 steps, batch 4, guidance=null (the
  reference's CPU-runnable case), full sample() wall time."""
  over = ['parameterization=d3pm', 'diffusion=uniform', 'time_conditioning=True', 'T=0',
          'model=tiny', 'model.length=256', 'training.guidance=null']
  for B in (4, 512):
    cfg = ddg_b200.compose(over + [f'sampling.batch_size={B}', 'sampling.steps=128'])
    tok = Text8Tok()
    sd = ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=1)
    model = ddg_b200.Diffusion(cfg, tok, device=dev).load_state_dict(sd)
    model.sample()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.sample()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert int(out.min()) >= 0 and int(out.max()) < tok.vocab_size
    print(f'{"C1 text8 UDLM tiny null B=%d" % B:32s} {dt / 128 * 1e3:8.3f} ms/step  -> {B / dt:7.2f} seq/s at 128 steps',
          flush=True)
    del model
    torch.cuda.empty_cache()


def main():
  dev = torch.device('cuda', 0)
  steps = 40
  c1_main(dev)
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
  cbg_main(dev)


if __name__ == '__main__':
  main()
