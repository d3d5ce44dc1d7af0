"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference DiT denoiser.

Plain torch-on-CPU fp32 restatement of `DIT.forward` / `DITClassifier.forward`
(reference models/dit.py:426-468, 516-556) operating directly on a reference
`state_dict` (key names of SURVEY.md section 8a).  It exists to check the CUDA
path; the product package never imports it.

Parity status: pinned against the unmodified reference imported in the build
container (tests/golden/make_golden.py writes fixtures, tests/test_oracle_pins.py
replays them; the `*_amp` fixtures pin `Rounding.gpu_*` against the reference run
under an emulation of its cuda autocast policy).  The reference itself ships no
tests or golden vectors.

`Rounding` reproduces where the reference rounds on a GPU (SURVEY.md 8a'):
blocks under autocast(fp16) (models/dit.py:454), final layer under
autocast(bf16) (:463), classifier blocks under autocast(bf16) (:531).  With
`Rounding.none()` everything is fp32, which is what the reference does on CPU
(torch disables cuda autocast there).
"""
from __future__ import annotations

import dataclasses
import math
import typing

import torch
import torch.nn.functional as F


@dataclasses.dataclass(frozen=True)
class Rounding:
  block: typing.Optional[torch.dtype] = None   # dtype of the block autocast
  head: typing.Optional[torch.dtype] = None    # dtype of the final-layer autocast

  @staticmethod
  def none():
    return Rounding(None, None)

  @staticmethod
  def gpu_denoiser():      # models/dit.py:454, 463
    return Rounding(torch.float16, torch.bfloat16)

  @staticmethod
  def gpu_classifier():    # models/dit.py:531, 538
    return Rounding(torch.bfloat16, torch.bfloat16)


def _r(x, dt):
  """Round-trip through `dt` (what an autocast op's output dtype does)."""
  return x if dt is None else x.to(dt).float()


def _linear(x, w, b, dt):
  """nn.Linear under autocast(dt): operands rounded to dt, fp32 accumulate,
  bias added before the single output rounding."""
  y = _r(x, dt) @ _r(w, dt).t()
  if b is not None:
    y = y + _r(b, dt)
  return _r(y, dt)


def timestep_embedding(t, dim=256, max_period=10000.0):
  # models/dit.py:169-193
  half = dim // 2
  freqs = torch.exp(
    -math.log(max_period)
    * torch.arange(half, dtype=torch.float32) / half)
  args = t[:, None].float() * freqs[None]
  return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def conditioning(sd, prefix, sigma, cond, causal=False):
  """c = silu(sigma_map(sigma)) [+ silu(cond_map(cond))]; fp32 everywhere
  (models/dit.py:433-446; outside every inner autocast)."""
  c = None
  if not causal:
    h = timestep_embedding(sigma)
    h = F.linear(h, sd[prefix + 'sigma_map.mlp.0.weight'],
                 sd[prefix + 'sigma_map.mlp.0.bias'])
    h = F.silu(h)
    h = F.linear(h, sd[prefix + 'sigma_map.mlp.2.weight'],
                 sd[prefix + 'sigma_map.mlp.2.bias'])
    c = F.silu(h)
  if cond is not None:
    key = prefix + 'cond_map.embedding_table.weight'
    if key not in sd:
      raise ValueError('Conditioning variable provided, but Model was not '
                       'initialized with condition embedding layer.')
    e = F.silu(sd[key][cond])
    c = e if c is None else c + e
  return c


def rotary_tables(seq_len, head_dim, base=10000.0):
  # models/dit.py:82-108 (inv_freq) -> cos/sin (L, head_dim/2)
  inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
  t = torch.arange(seq_len, dtype=torch.float32)
  freqs = torch.einsum('i,j->ij', t, inv_freq)
  return freqs.cos(), freqs.sin()


def _rope(x, cos, sin, dt):
  """NeoX (non-interleaved) rotary on (B, L, H, D); cos/sin rounded to the
  qkv dtype first, math in fp32, result rounded once
  (models/dit.py:290-293; flash_attn rotary kernel)."""
  d2 = x.shape[-1] // 2
  c = _r(cos, dt)[None, :, None, :]
  s = _r(sin, dt)[None, :, None, :]
  x1, x2 = x[..., :d2], x[..., d2:]
  return _r(torch.cat([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1), dt)


def _attention(q, k, v, causal, dt):
  """SDPA (models/dit.py:303-307) the way flash kernels evaluate it: fp32
  scores, unnormalised exp rounded to dt before P.V, normalise at the end."""
  scale = 1.0 / math.sqrt(q.shape[-1])
  s = torch.einsum('blhd,bmhd->bhlm', q, k) * scale
  if causal:
    L = s.shape[-1]
    mask = torch.ones(L, L, dtype=torch.bool).tril()
    s = s.masked_fill(~mask, float('-inf'))
  m = s.amax(dim=-1, keepdim=True)
  p = torch.exp(s - m)
  l = p.sum(dim=-1, keepdim=True)
  o = torch.einsum('bhlm,bmhd->bhld', _r(p, dt), v) / l
  return _r(o.permute(0, 2, 1, 3), dt)       # (B, L, H, D)


def _layer_norm(x, w):
  # models/dit.py:131-139: fp32 stats, no bias, then * weight
  return F.layer_norm(x.float(), [x.shape[-1]]) * w[None, None, :]


def _modulate(x, shift, scale, dt):
  # effective definition models/dit.py:47-50 (scripted :75-79); (1+scale) is
  # formed in the autocast dtype, the product/sum promote to fp32.
  return x * _r(1.0 + scale, dt) + shift


def _block(sd, p, x, c, cos, sin, n_heads, causal, use_adaln, dt):
  B, L, H = x.shape
  if use_adaln:
    mod = _linear(c, sd[p + 'adaLN_modulation.weight'],
                  sd[p + 'adaLN_modulation.bias'], dt)[:, None]
    (sh_a, sc_a, g_a, sh_m, sc_m, g_m) = mod.chunk(6, dim=2)
  # attention half (models/dit.py:277-316)
  skip = x
  h = _layer_norm(x, sd[p + 'norm1.weight'])
  if use_adaln:
    h = _modulate(h, sh_a, sc_a, dt)
  qkv = _linear(h, sd[p + 'attn_qkv.weight'], None, dt)
  qkv = qkv.view(B, L, 3, n_heads, H // n_heads)
  q = _rope(qkv[:, :, 0], cos, sin, dt)
  k = _rope(qkv[:, :, 1], cos, sin, dt)
  v = qkv[:, :, 2]
  a = _attention(q, k, v, causal, dt).reshape(B, L, H)
  o = _linear(a, sd[p + 'attn_out.weight'], None, dt)
  if use_adaln:
    o = _r(g_a * o, dt)
  x = skip + o
  # mlp half (models/dit.py:318-325)
  skip = x
  h = _layer_norm(x, sd[p + 'norm2.weight'])
  if use_adaln:
    h = _modulate(h, sh_m, sc_m, dt)
  h = _linear(h, sd[p + 'mlp.0.weight'], sd[p + 'mlp.0.bias'], dt)
  h = _r(F.gelu(h, approximate='tanh'), dt)
  h = _linear(h, sd[p + 'mlp.2.weight'], sd[p + 'mlp.2.bias'], dt)
  if use_adaln:
    h = _r(g_m * h, dt)
  return skip + h


def _n_blocks(sd, prefix):
  n = 0
  while f'{prefix}blocks.{n}.norm1.weight' in sd:
    n += 1
  return n


def dit_forward(sd, ids, sigma, cond=None, *, n_heads, causal=False,
                prefix='backbone.', rounding=Rounding.none(),
                return_hidden=False):
  """DIT.forward (models/dit.py:426-468): ids (B,L) int64, sigma (B,) f32 (or
  None when causal), cond (B,) int64 or None -> logits (B,L,V) fp32 holding
  values representable in the head dtype."""
  c = conditioning(sd, prefix, sigma, cond, causal)
  use_adaln = (prefix + 'blocks.0.adaLN_modulation.weight') in sd
  x = sd[prefix + 'vocab_embed.embedding'][ids]
  B, L, H = x.shape
  cos, sin = rotary_tables(L, H // n_heads)
  for i in range(_n_blocks(sd, prefix)):
    x = _block(sd, f'{prefix}blocks.{i}.', x, c, cos, sin, n_heads, causal,
               use_adaln, rounding.block)
  hidden = x
  dt = rounding.head
  p = prefix + 'output_layer.'
  h = _layer_norm(x, sd[p + 'norm_final.weight'])
  if (p + 'adaLN_modulation.weight') in sd:
    mod = _linear(c, sd[p + 'adaLN_modulation.weight'],
                  sd[p + 'adaLN_modulation.bias'], dt)[:, None]
    shift, scale = mod.chunk(2, dim=2)
    h = _modulate(h, shift, scale, dt)
  logits = _linear(h, sd[p + 'linear.weight'], sd[p + 'linear.bias'], dt)
  if return_hidden:
    return logits, hidden
  return logits


def classifier_forward(sd, ids_or_one_hot, sigma, *, n_heads, causal=False,
                       prefix='classifier_model.',
                       rounding=Rounding.none()):
  """DITClassifier.forward with mean pooling (models/dit.py:516-556):
  -> class logits (n, num_classes)."""
  E = sd[prefix + 'vocab_embed.embedding']
  if ids_or_one_hot.ndim == 2:
    x = E[ids_or_one_hot]
  else:
    x = ids_or_one_hot.float() @ E      # F.linear(one_hot, E.T)
  c = conditioning(sd, prefix, sigma, None, causal)
  B, L, H = x.shape
  cos, sin = rotary_tables(L, H // n_heads)
  for i in range(_n_blocks(sd, prefix)):
    x = _block(sd, f'{prefix}blocks.{i}.', x, c, cos, sin, n_heads, causal,
               not causal, rounding.block)
  x = x.mean(dim=1)
  return _linear(x, sd[prefix + 'output_layer.weight'],
                 sd[prefix + 'output_layer.bias'], rounding.head)


def classifier_log_probs(sd, ids_or_one_hot, sigma, *, n_heads,
                         time_conditioning, prefix='classifier_model.',
                         rounding=Rounding.none()):
  """Classifier.get_log_probs (reference classifier.py:316-327, 484-490)."""
  sigma = sigma.reshape(-1)
  if not time_conditioning:
    sigma = torch.zeros_like(sigma)
  logits = classifier_forward(sd, ids_or_one_hot, sigma, n_heads=n_heads,
                              prefix=prefix, rounding=rounding)
  # fp32 on the GPU as well: the log_softmax sits inside `autocast(dtype=float32)`
  # (classifier.py:325-327), which CUDA autocast does NOT disable -- log_softmax is on
  # its fp32 list, so the bf16 class logits are cast up.  Measured on the B200
  # (classifier_log_prob_dtype == float32 in tests/golden/*_cbg_gpu.npz); SURVEY 8a'
  # item 6 assumed bf16.
  return F.log_softmax(logits, dim=-1)
