"""Golden-vector case definitions shared by the generator (make_golden.py, run
in the build container against the unmodified reference) and the tests that
replay them (oracle on CPU, CUDA path on the GPU box).

Everything random is derived from seeds so fixtures stay small: weights come
from `ddg_b200.weights.random_state_dict(seed)`, the prior from
`prior_ids(case)`, step i's uniforms from `step_noise(case, i)`.
"""
from __future__ import annotations

import dataclasses
import typing

import torch


@dataclasses.dataclass(frozen=True)
class Case:
  name: str
  vocab: str = 'text8'          # 'text8' (V=35, [MASK]=4) | 'bert' (V=30522, [MASK]=103)
  parameterization: str = 'd3pm'
  diffusion: str = 'uniform'
  time_conditioning: bool = True
  subs_masking: bool = False
  T: int = 0
  guidance: typing.Optional[str] = None   # None | 'cfg' | 'cbg'
  gamma: float = 1.0
  condition: int = 1
  num_classes: int = 2
  use_approx: bool = False
  argmax_sampling: bool = False
  use_float64: bool = False
  low_confidence: bool = False          # eval.low_confidence_sampling
  low_confidence_threshold: float = 0.3
  use_cache: bool = False               # sampling.use_cache (diffusion.py:1371-1374)
  batch: int = 2
  length: int = 32
  steps: int = 4
  seed: int = 11
  hidden: int = 128
  n_heads: int = 2
  n_blocks: int = 2
  cond_dim: int = 64


CASES = [
  Case('udlm_null'),
  Case('mdlm_null', parameterization='subs', diffusion='absorbing_state',
       time_conditioning=False, steps=6),
  Case('udlm_cfg_g2', guidance='cfg', gamma=2.0),
  Case('udlm_cfg_g0', guidance='cfg', gamma=0.0, steps=2),
  Case('udlm_cfg_g1', guidance='cfg', gamma=1.0, steps=2),
  Case('mdlm_cfg_g2', parameterization='subs', diffusion='absorbing_state',
       time_conditioning=False, guidance='cfg', gamma=2.0, steps=5),
  Case('udlm_cbg', guidance='cbg', gamma=1.5, length=16, steps=2),
  Case('mdlm_cbg', parameterization='subs', diffusion='absorbing_state',
       time_conditioning=False, guidance='cbg', gamma=1.5, length=16, steps=3),
  Case('udlm_cbg_approx', guidance='cbg', gamma=1.5, use_approx=True, length=16,
       steps=2),
  Case('mdlm_cbg_approx', parameterization='subs', diffusion='absorbing_state',
       time_conditioning=False, guidance='cbg', gamma=1.5, use_approx=True,
       length=16, steps=3),
  Case('udlm_T1000', T=1000, steps=3),
  Case('d3pm_abs_submask', diffusion='absorbing_state', subs_masking=True,
       steps=4),
  Case('udlm_argmax', argmax_sampling=True, steps=2),
  Case('bert_udlm_cfg_g2', vocab='bert', guidance='cfg', gamma=2.0, length=16,
       steps=2),
  Case('bert_mdlm_null', vocab='bert', parameterization='subs',
       diffusion='absorbing_state', time_conditioning=False, length=16,
       steps=3),
  Case('udlm_f64', use_float64=True, steps=3),
  Case('mdlm_f64', parameterization='subs', diffusion='absorbing_state',
       time_conditioning=False, use_float64=True, steps=4),
  Case('udlm_lowconf', low_confidence=True, steps=3),
  Case('mdlm_lowconf', parameterization='subs', diffusion='absorbing_state',
       time_conditioning=False, low_confidence=True,
       low_confidence_threshold=0.6, steps=4),
  Case('udlm_cfg_lowconf', guidance='cfg', gamma=2.0, low_confidence=True,
       low_confidence_threshold=0.5, steps=2),
  Case('bert_udlm_lowconf', vocab='bert', low_confidence=True, length=16,
       steps=2),
  Case('udlm_use_cache', use_cache=True, batch=1, length=16, steps=24),
  Case('udlm_cfg_use_cache', use_cache=True, guidance='cfg', gamma=2.0, batch=1,
       length=16, steps=24),
  Case('ar_null', parameterization='ar', diffusion='absorbing_state',
       time_conditioning=False, length=16, steps=0),
  Case('ar_cfg_g2', parameterization='ar', diffusion='absorbing_state',
       time_conditioning=False, guidance='cfg', gamma=2.0, length=16, steps=0),
  Case('ar_lowconf', parameterization='ar', diffusion='absorbing_state',
       time_conditioning=False, low_confidence=True,
       low_confidence_threshold=0.6, length=16, steps=0),
]
CASE_BY_NAME = {c.name: c for c in CASES}
# cases also run with the reference's GPU autocast policy emulated on the CPU
# (make_golden.py --amp -> <name>_amp.npz): they pin the rounding points
AMP_CASES = ['udlm_null', 'mdlm_null', 'udlm_cfg_g2', 'mdlm_cfg_g2', 'udlm_cbg',
             'mdlm_cbg', 'bert_udlm_cfg_g2', 'bert_mdlm_null', 'udlm_cbg_approx',
             'mdlm_cbg_approx']


class Tok:
  """Tokenizer surrogate: the sampling path only reads these attributes
  (reference diffusion.py:80-95, 1093)."""

  def __init__(self, vocab):
    if vocab == 'text8':    # dataloader.py:49-101: 35 ids, [MASK]=4, [BOS]=2
      self.vocab_size, self.mask_token_id = 35, 4
      self.bos_token_id, self.eos_token_id, self.pad_token_id = 2, 3, 5
    else:                   # bert-base-uncased
      self.vocab_size, self.mask_token_id = 30522, 103
      self.bos_token_id, self.eos_token_id, self.pad_token_id = 101, 102, 0
    self.mask_token = '[MASK]'


def config_dict(case: Case) -> dict:
  """Config with the reference's key names as a plain dict."""
  model = dict(name='test', type='ddit', hidden_size=case.hidden,
               cond_dim=case.cond_dim, length=case.length,
               n_blocks=case.n_blocks, n_heads=case.n_heads,
               scale_by_sigma=True, dropout=0.1, tie_word_embeddings=False)
  guidance = None
  if case.guidance == 'cfg':
    guidance = dict(method='cfg', condition=case.condition, gamma=case.gamma)
  elif case.guidance == 'cbg':
    guidance = dict(method='cbg', condition=case.condition, gamma=case.gamma,
                    classifier_checkpoint_path='<injected>',
                    use_approx=case.use_approx)
  cfg = dict(
    mode='sample', diffusion=case.diffusion, backbone='dit',
    classifier_backbone='dit' if case.guidance == 'cbg' else None,
    parameterization=case.parameterization,
    time_conditioning=case.time_conditioning, subs_masking=case.subs_masking,
    zero_recon_loss=False, T=case.T, is_vision=False, seed=case.seed,
    model=model,
    classifier_model=(dict(model, n_blocks=1, num_classes=case.num_classes,
                           pooling='mean')
                      if case.guidance == 'cbg' else None),
    guidance=guidance,
    noise=dict(type='loglinear', sigma_min=1e-4, sigma_max=20),
    data=dict(train='synthetic', valid='synthetic',
              num_classes=case.num_classes),
    sampling=dict(use_cache=case.use_cache, steps=max(case.steps, 1),
                  batch_size=case.batch, num_sample_batches=1,
                  use_float64=case.use_float64),
    training=dict(guidance=(dict(cond_dropout=0.1)
                            if case.guidance == 'cfg' else None),
                  antithetic_sampling=True, importance_sampling=False,
                  sampling_eps=1e-3, change_of_variables=False,
                  label_smoothing=0.0),
    eval=dict(checkpoint_path='', disable_ema=True,
              low_confidence_sampling=case.low_confidence,
              low_confidence_threshold=case.low_confidence_threshold,
              argmax_sampling=case.argmax_sampling),
    optim=dict(lr=3e-4, weight_decay=0, beta1=0.9, beta2=0.999, eps=1e-8),
    weights_averaging=dict(type='none'),
    trainer=dict(max_steps=1),
  )
  return cfg


def vocab_size(case: Case) -> int:
  return Tok(case.vocab).vocab_size


def prior_ids(case: Case) -> torch.Tensor:
  g = torch.Generator().manual_seed(case.seed * 7 + 1)
  tok = Tok(case.vocab)
  if case.parameterization == 'ar':
    return torch.zeros(case.batch, case.length, dtype=torch.int64)
  if case.diffusion == 'absorbing_state':
    return torch.full((case.batch, case.length), tok.mask_token_id,
                      dtype=torch.int64)
  return torch.randint(0, tok.vocab_size, (case.batch, case.length),
                       generator=g, dtype=torch.int64)


def step_noise(case: Case, i: int) -> torch.Tensor:
  """Uniform [0,1) fp32 (B, L, V) for step i (what torch.rand_like would have
  drawn in `_sample_categorical`, diffusion.py:1752-1754)."""
  g = torch.Generator().manual_seed(case.seed * 1000 + i)
  return torch.rand(case.batch, case.length, vocab_size(case), generator=g,
                    dtype=torch.float32)


def ar_gumbel(case: Case) -> torch.Tensor:
  """Gumbel(0,1) noise (B, L-1, V) for the AR sampler (diffusion.py:1106)."""
  g = torch.Generator().manual_seed(case.seed * 1000 + 999)
  u = torch.rand(case.batch, case.length - 1, vocab_size(case), generator=g,
                 dtype=torch.float32).clamp_(1e-9, 1 - 1e-7)
  return -torch.log(-torch.log(u))
