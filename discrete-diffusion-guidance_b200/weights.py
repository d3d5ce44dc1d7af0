"""Weight ingestion for the sampling path.

`state_dict` key names and shapes are the reference's (`models/dit.py:369-418,
470-514`; listed in SURVEY.md 8a), so a Lightning checkpoint of the reference
loads without renaming.  `random_state_dict` builds a seeded random-init of
the same architecture for benchmarks/tests (no checkpoints are reachable
offline); unlike the reference's init it does not leave the head and adaLN
tensors at zero (`models/dit.py:253-254, 347-355`), which would make every
logit 0.
"""
from __future__ import annotations

import math
import typing

import torch


def has_cond_map(config) -> bool:
  # models/dit.py:387-395
  tg = config.get('training') and config.training.get('guidance')
  g = config.get('guidance')
  return tg is not None or (g is not None and g.get('method') == 'cfg')


def random_state_dict(config, vocab_size: int, seed: int = 0,
                      classifier: bool = False, std: float = 0.02,
                      head_std: float = 0.08) -> typing.Dict[str, torch.Tensor]:
  """Seeded fp32 state_dict with the reference's names ('backbone.*' or
  'classifier_model.*')."""
  g = torch.Generator(device='cpu')
  g.manual_seed(seed)
  m = config.classifier_model if classifier else config.model
  h, cd, nb, nh = m.hidden_size, m.cond_dim, m.n_blocks, m.n_heads
  causal = config.parameterization == 'ar'
  prefix = 'classifier_model.' if classifier else 'backbone.'
  sd = {}

  def normal(*shape, s=std):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * s

  def near_one(n):
    return 1.0 + 0.1 * torch.randn(n, generator=g, dtype=torch.float32)

  bound = 1.0 / math.sqrt(h)
  sd['vocab_embed.embedding'] = (
    torch.rand(vocab_size, h, generator=g, dtype=torch.float32) * 2 - 1) * bound * 4
  if not causal:
    sd['sigma_map.mlp.0.weight'] = normal(cd, 256, s=0.05)
    sd['sigma_map.mlp.0.bias'] = normal(cd, s=0.05)
    sd['sigma_map.mlp.2.weight'] = normal(cd, cd, s=0.08)
    sd['sigma_map.mlp.2.bias'] = normal(cd, s=0.05)
  cond_map = (not classifier) and has_cond_map(config)
  if cond_map:
    sd['cond_map.embedding_table.weight'] = normal(
      config.data.num_classes + 1, cd, s=1.0)
  hd = h // nh
  sd['rotary_emb.inv_freq'] = 1.0 / (
    10000 ** (torch.arange(0, hd, 2).float() / hd))
  use_adaln = (not causal) if classifier else ((not causal) or cond_map)
  for i in range(nb):
    p = f'blocks.{i}.'
    sd[p + 'norm1.weight'] = near_one(h)
    sd[p + 'attn_qkv.weight'] = normal(3 * h, h, s=0.04)
    sd[p + 'attn_out.weight'] = normal(h, h)
    sd[p + 'norm2.weight'] = near_one(h)
    sd[p + 'mlp.0.weight'] = normal(4 * h, h)
    sd[p + 'mlp.0.bias'] = normal(4 * h)
    sd[p + 'mlp.2.weight'] = normal(h, 4 * h)
    sd[p + 'mlp.2.bias'] = normal(h)
    if use_adaln:
      sd[p + 'adaLN_modulation.weight'] = normal(6 * h, cd, s=0.05)
      sd[p + 'adaLN_modulation.bias'] = normal(6 * h, s=0.05)
  if classifier:
    sd['output_layer.weight'] = normal(m.num_classes, h, s=0.5)
    sd['output_layer.bias'] = normal(m.num_classes, s=0.1)
  else:
    sd['output_layer.norm_final.weight'] = near_one(h)
    sd['output_layer.linear.weight'] = normal(vocab_size, h, s=head_std)
    sd['output_layer.linear.bias'] = normal(vocab_size, s=0.5)
    if use_adaln:
      sd['output_layer.adaLN_modulation.weight'] = normal(2 * h, cd, s=0.05)
      sd['output_layer.adaLN_modulation.bias'] = normal(2 * h, s=0.05)
  return {prefix + k: v for k, v in sd.items()}


def load_checkpoint_state(path: str, use_ema: bool = True,
                          prefix: str = 'backbone.'):
  """Reads a Lightning .ckpt of the reference (diffusion.py:194-213): returns
  the `prefix`-ed tensors, with the EMA shadow parameters substituted when
  present (`['ema']['shadow_params']` is ordered like
  `chain(backbone.parameters(), noise.parameters())`, diffusion.py:144-151)."""
  ckpt = torch.load(path, map_location='cpu', weights_only=False)
  sd = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
  out = {k: v.float() for k, v in sd.items() if k.startswith(prefix)}
  ema = ckpt.get('ema') if isinstance(ckpt, dict) else None
  if use_ema and ema and ema.get('shadow_params'):
    # parameters() order == state_dict order minus buffers (rotary inv_freq)
    names = [k for k in out if not k.endswith('rotary_emb.inv_freq')]
    shadow = ema['shadow_params']
    for k, v in zip(names, shadow):
      if tuple(v.shape) != tuple(out[k].shape):
        raise ValueError(f'EMA shadow param for {k}: shape {tuple(v.shape)} '
                         f'!= {tuple(out[k].shape)}')
      out[k] = v.float()
  return out
