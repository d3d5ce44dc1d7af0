"""Guidance classifier for D-CBG: mirror of `classifier.Classifier`'s inference
surface (reference classifier.py:132-216, 300-327, 484-490) over the
`DITClassifier` backbone (models/dit.py:470-556)."""
from __future__ import annotations

import torch

from . import weights
from .dit import DIT


class Classifier:

  def __init__(self, config, tokenizer, device=None):
    self.config = config
    self.tokenizer = tokenizer
    # classifier.py:151-156
    self.vocab_size = tokenizer.vocab_size
    if (not hasattr(tokenizer, 'mask_token') or tokenizer.mask_token is None):
      self.mask_index = self.vocab_size
      self.vocab_size += 1
    else:
      self.mask_index = tokenizer.mask_token_id
    if config.classifier_backbone != 'dit':
      raise NotImplementedError(
        f'Classifier backbone {config.classifier_backbone} not implemented.')
    self.classifier_model = DIT(config, vocab_size=self.vocab_size,
                                device=device, classifier=True)
    self.time_conditioning = config.time_conditioning
    self.round_bf16 = True
    self.chunk_sequences = 16384   # candidate sequences per classifier launch

  def load_state_dict(self, state_dict, strict=True):
    self.classifier_model.load_state_dict(state_dict,
                                          prefix='classifier_model.')
    return self

  @classmethod
  def load_from_checkpoint(cls, checkpoint_path, tokenizer, config,
                           device=None, **_):
    self = cls(config, tokenizer, device=device)
    sd = weights.load_checkpoint_state(checkpoint_path, use_ema=False,
                                       prefix='classifier_model.')
    return self.load_state_dict(sd)

  def eval(self):
    return self

  def to(self, device):
    return self

  def _process_sigma(self, sigma):
    # classifier.py:484-490
    if sigma.ndim > 1:
      sigma = sigma.squeeze(-1)
    if not self.time_conditioning:   # one shared zero (stride 0): a single conditioning row
      sigma = torch.zeros(1, dtype=sigma.dtype, device=sigma.device).expand(sigma.shape[0])
    assert sigma.ndim == 1, sigma.shape
    return sigma

  def get_log_probs(self, x, sigma, x_emb=None, seqs_per_mod=1):
    """classifier.py:316-327 -> (n, classes) fp32 log-probs.  `x` is (n, L) token
    ids or, like `DITClassifier.forward` (models/dit.py:518-522), (n, L, V)
    one-hots: `one_hot @ embedding` is the embedding gather, so one-hot rows are
    turned back into ids (rows that are not one-hot raise ValueError; the gradient
    w.r.t. a one-hot input is `first_order_term`).
    `seqs_per_mod` > 1: sigma holds one value per that many consecutive sequences."""
    if x_emb is not None:
      raise NotImplementedError(
        'x_emb inputs (NOS / PPLM hidden-state guidance) are outside the sampling path')
    if x.ndim == 3:
      if x.shape[-1] != self.vocab_size:
        raise ValueError(f'one-hot input has {x.shape[-1]} columns, vocab is {self.vocab_size}')
      top, ids = x.max(dim=-1)
      if not bool(((top == 1) & (x.sum(dim=-1) == 1) & (x.min(dim=-1).values == 0)).all()):
        raise ValueError('classifier input of shape (n, L, V) must be one-hot rows')
      x = ids
    elif x.ndim != 2:
      raise ValueError(f'classifier input must be (n, L) ids or (n, L, V) one-hots, got {tuple(x.shape)}')
    return self.classifier_model.classify(x, self._process_sigma(sigma),
                                          round_bf16=self.round_bf16,
                                          seqs_per_mod=seqs_per_mod)

  def first_order_term(self, xt, sigma, conditioning_class, finalize=True, out=None, live=None):
    """The `use_approx` branch of `_cbg_denoise` (diffusion.py:1532-1548):
    `grad - (one_hot * grad).sum(-1) + log p(y | xt)` with `grad` the gradient
    of `get_log_probs(one_hot(xt), sigma)[..., y].sum()` w.r.t. the one-hot
    input.  Returns ((B, L, V) fp32 term, (B, classes) log-probs)."""
    return self.classifier_model.first_order_term(
      xt, self._process_sigma(sigma), conditioning_class,
      round_bf16=self.round_bf16, finalize=finalize, out=out, live=live)
