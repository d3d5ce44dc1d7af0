"""TEST / BENCH INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference `Diffusion`
(and `Classifier`) through oracle/refshim.py with the same seeded random-init
weights the CUDA path is given (`weights.random_state_dict`), on the CPU or on a
GPU.  Used by tests/golden/make_golden.py, tests/test_gpu_reference.py and
`bench.py --impl reference`.

The product's `config.py` / `weights.py` are pure Python + torch; they are
loaded here by file path, NOT through the `ddg_b200` package, so that a process
that only drives the reference never maps libddg_b200.so.
"""
from __future__ import annotations

import importlib.util
import os
import sys

from oracle import refshim

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    'discrete-diffusion-guidance_b200')
_HOST = {}


def host_module(name: str):
  """`config` / `weights` of the product package without importing the package
  (whose __init__ loads the CUDA library)."""
  if name not in _HOST:
    full = 'ddg_b200.' + name
    if full in sys.modules:          # package already imported (tests): reuse
      _HOST[name] = sys.modules[full]
    else:
      spec = importlib.util.spec_from_file_location(
        f'_ddg_host_{name}', os.path.join(_PKG, name + '.py'))
      mod = importlib.util.module_from_spec(spec)
      spec.loader.exec_module(mod)
      _HOST[name] = mod
  return _HOST[name]


def random_weights(cfg_dict: dict, vocab_size: int, seed: int, classifier=False):
  cfg = host_module('config').to_config(cfg_dict)
  return host_module('weights').random_state_dict(cfg, vocab_size, seed=seed,
                                                  classifier=classifier)


def build_reference(cfg_dict: dict, tok, seed: int, device='cpu',
                    with_classifier: bool = False):
  """-> (reference Diffusion, reference Classifier | None, cfg, state_dict,
  classifier state_dict | None), all in eval mode on `device`."""
  ref = refshim.load_reference()
  cfg = refshim.to_attrdict(cfg_dict)
  model = ref.diffusion.Diffusion(cfg, tok)
  sd = random_weights(cfg_dict, tok.vocab_size, seed)
  missing, unexpected = model.load_state_dict(sd, strict=False)
  assert not unexpected, unexpected
  assert set(missing) <= {'limiting_distribution'}, missing
  model.eval().to(device)
  clf = csd = None
  if with_classifier:
    clf = ref.classifier.Classifier(cfg, tok)
    csd = random_weights(cfg_dict, tok.vocab_size, seed + 1, classifier=True)
    missing, unexpected = clf.load_state_dict(csd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    clf.eval().to(device)
  return model, clf, cfg, sd, csd


class BertTok:
  """bert-base-uncased surrogate (no network): the sampling path reads only these
  attributes (reference diffusion.py:80-95)."""
  vocab_size, mask_token, mask_token_id = 30522, '[MASK]', 103
  bos_token_id, eos_token_id, pad_token_id = 101, 102, 0


def _cases():
  here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  spec = importlib.util.spec_from_file_location(
    '_ddg_golden_cases', os.path.join(here, 'tests', 'golden', 'cases.py'))
  mod = importlib.util.module_from_spec(spec)
  sys.modules[spec.name] = mod
  spec.loader.exec_module(mod)
  return mod


def config3_dict(batch: int, steps: int = 1000, gamma: float = 2.0) -> dict:
  """BASELINE.json configs[2] in the reference's key names: LM1B UDLM (d3pm,
  uniform) DiT-small (configs/model/small.yaml), L=128, V=30522, D-CFG."""
  cases = _cases()
  case = cases.Case('config3', vocab='bert', parameterization='d3pm',
                    diffusion='uniform', time_conditioning=True, guidance='cfg',
                    gamma=gamma, condition=1, num_classes=2, batch=batch,
                    length=128, steps=steps, seed=1, hidden=768, n_heads=12,
                    n_blocks=12, cond_dim=128)
  return cases.config_dict(case)


class ReferenceStepper:
  """Drives the reference's per-step update the way `_diffusion_sample`
  (reference diffusion.py:1293-1366) does, one call per reverse step, so a
  benchmark can time exactly K steps."""

  def __init__(self, model, cfg, batch, device='cpu', eps=1e-5):
    import torch
    self.torch = torch
    self.model, self.cfg, self.B, self.device = model, cfg, batch, device
    steps = cfg.sampling.steps
    self.timesteps = torch.linspace(1, eps, steps + 1, device=device)
    self.dt = (1 - eps) / steps
    self.cond = (torch.ones(batch, device=device) * cfg.guidance.condition).to(torch.long) \
        if cfg.guidance is not None and cfg.guidance.method == 'cfg' else None

  def step(self, xt, i):
    torch, m = self.torch, self.model
    t = self.timesteps[i % (self.timesteps.numel() - 1)] * torch.ones(
      self.B, 1, device=self.device)
    sigma_t, _ = m.noise(t)
    sigma_s, _ = m.noise(t - self.dt)
    sigma_t, sigma_s = sigma_t.squeeze(-1), sigma_s.squeeze(-1)
    mct = (1 - torch.exp(-sigma_t))[:, None, None]
    mcs = (1 - torch.exp(-sigma_s))[:, None, None]
    with torch.no_grad():
      if self.cond is None:
        xs, _, _ = m._ddpm_denoise(xt=xt, time_conditioning=sigma_t,
                                   move_chance_t=mct, move_chance_s=mcs, cache=None)
      else:
        xs, _, _ = m._cfg_denoise(cond=self.cond, gamma=self.cfg.guidance.gamma, xt=xt,
                                  time_conditioning=sigma_t, move_chance_t=mct,
                                  move_chance_s=mcs, cache=None)
    return xs


def make_reference_checkpoints(cfg_dict: dict, tok, seed: int, out_dir: str,
                               with_classifier: bool = True):
  """Writes Lightning-format `.ckpt` files from the REFERENCE modules the way
  `on_save_checkpoint` (reference diffusion.py:208-250) shapes them:
  `{'state_dict': ..., 'ema': ema.state_dict(), 'loops': ...}` with the EMA shadow
  parameters perturbed away from the raw ones, and a classifier checkpoint
  (`classifier_model.*`).  Also returns what the reference itself would sample
  with: the state_dict after its own `ema.copy_to(...)` (diffusion.py:923-930).

  -> dict(diffusion=path, classifier=path|None, raw_sd, ema_sd, clf_sd)"""
  import copy
  import itertools

  import torch
  cfg_dict = copy.deepcopy(cfg_dict)
  cfg_dict['weights_averaging'] = dict(type='ema', decay=0.9999)
  model, clf, cfg, sd, csd = build_reference(cfg_dict, tok, seed,
                                             with_classifier=with_classifier)
  assert model.ema is not None
  g = torch.Generator().manual_seed(seed + 77)
  for p in model.ema.shadow_params:          # an EMA that differs from the raw weights
    p.add_(torch.randn(p.shape, generator=g) * 0.01 * (p.abs().mean() + 0.01))
  raw_sd = {k: v.clone() for k, v in model.state_dict().items()
            if k != 'limiting_distribution'}
  ckpt = {'state_dict': raw_sd, 'ema': model.ema.state_dict(),
          'loops': {'fit_loop': {'epoch_progress': {'current': {'completed': 0}},
                                 'epoch_loop.batch_progress': {
                                   'current': {'completed': 0},
                                   'total': {'completed': 0}}}},
          'pytorch-lightning_version': 'stub', 'global_step': 0, 'epoch': 0}
  os.makedirs(out_dir, exist_ok=True)
  dpath = os.path.join(out_dir, 'diffusion.ckpt')
  torch.save(ckpt, dpath)
  # what the reference samples with (diffusion.py:1054-1055): its own copy_to
  model.ema.copy_to(itertools.chain(model.backbone.parameters(),
                                    model.noise.parameters()))
  ema_sd = {k: v.clone() for k, v in model.state_dict().items()
            if k != 'limiting_distribution'}
  cpath = None
  if clf is not None:
    cpath = os.path.join(out_dir, 'classifier.ckpt')
    torch.save({'state_dict': {k: v.clone() for k, v in clf.state_dict().items()}}, cpath)
  return dict(diffusion=dpath, classifier=cpath, raw_sd=raw_sd, ema_sd=ema_sd,
              clf_sd=csd)
