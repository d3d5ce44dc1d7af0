"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference sampler math.

Plain torch-on-CPU restatement of the reverse-diffusion update of the
reference (`diffusion.py:1281-1624, 1738-1755`, `noise_schedule.py:134-160`),
written from the formulas, with the uniform noise `U` injected instead of
drawn.  It exists to check the CUDA path (tests/, smoke(), bench.py's CPU
baseline legs); the product package never imports it.

Parity status: pinned against the unmodified reference imported in the build
container -- see tests/golden/make_golden.py and tests/test_oracle_pins.py:
22 fixtures of the reference on CPU (fp32 math; ids exact, probabilities 2e-4)
and 10 `*_amp` fixtures of the reference with its cuda autocast policy emulated
on the CPU (the GPU rounding points, `lp_dtype` / `clf_dtype` here).
The reference ships no tests or golden vectors of its own.
"""
from __future__ import annotations

import typing

import torch
import torch.nn.functional as F

from . import ref_dit

NEG_INF = -1_000_000.0   # diffusion.py:141


# --------------------------------------------------------------------------
# noise schedule + time grid
# --------------------------------------------------------------------------
def total_noise(t, noise_type='loglinear', eps=1e-3):
  """sigma(t). loglinear: noise_schedule.py:152-153; cosine :75-77;
  cosinesqr :92-94."""
  if noise_type == 'loglinear':
    return -torch.log1p(-(1 - eps) * t)
  if noise_type == 'cosine':
    return -torch.log(eps + (1 - eps) * torch.cos(t * torch.pi / 2))
  if noise_type == 'cosinesqr':
    return -torch.log(eps + (1 - eps) * torch.cos(t * torch.pi / 2) ** 2)
  raise NotImplementedError(noise_type)


def step_scalars(steps, eps=1e-5, T=0, noise_type='loglinear'):
  """Per-step (sigma_t, move_chance_t, move_chance_s) exactly as
  `_diffusion_sample` forms them (diffusion.py:1293-1323): fp32 tensors of
  shape (steps,)."""
  timesteps = torch.linspace(1, eps, steps + 1)
  dt = (1 - eps) / steps
  sig_t, mc_t, mc_s = [], [], []
  for i in range(steps):
    t = timesteps[i]
    if T > 0:
      t = (t * T).to(torch.int)
      t = t / T
      t += (1 / T)
    t = t * torch.ones(1, 1)
    sigma_t = total_noise(t, noise_type).squeeze(-1)
    sigma_s = total_noise(t - dt, noise_type).squeeze(-1)
    sig_t.append(sigma_t)
    mc_t.append(1 - torch.exp(-sigma_t))
    mc_s.append(1 - torch.exp(-sigma_s))
  return torch.cat(sig_t), torch.cat(mc_t), torch.cat(mc_s)


# --------------------------------------------------------------------------
# forward wrapper: logits -> log-probs
# --------------------------------------------------------------------------
def log_probs_from_logits(logits, xt, *, parameterization, mask_index,
                          subs_masking=False, lp_dtype=None):
  """Diffusion.forward after the backbone (diffusion.py:356-365) incl.
  `_subs_parameterization` (:321-337).  `lp_dtype=torch.bfloat16` reproduces
  the GPU behaviour where log_softmax runs on (and returns) bf16."""
  logits = logits.clone().float()
  if parameterization == 'subs':
    logits[..., mask_index] += NEG_INF
    if lp_dtype is not None:
      logits = logits.to(lp_dtype).float()
    unmasked = xt != mask_index
    logits[unmasked] = NEG_INF
    logits[unmasked, xt[unmasked]] = 0
  elif parameterization in ('d3pm', 'ar'):
    if subs_masking:
      logits[..., mask_index] += NEG_INF
      if lp_dtype is not None:
        logits = logits.to(lp_dtype).float()
  else:
    raise NotImplementedError(parameterization)
  lp = F.log_softmax(logits, dim=-1)
  return lp if lp_dtype is None else lp.to(lp_dtype).float()


def _exp(lp, lp_dtype):
  x = lp.exp()
  return x if lp_dtype is None else x.to(lp_dtype).float()


# --------------------------------------------------------------------------
# posteriors
# --------------------------------------------------------------------------
def posterior_uniform(x, xt, alpha_s, alpha_t):
  """UDLM q(z_s | z_t, x) (diffusion.py:367-394).  x (B,L,V), xt (B,L),
  alphas broadcastable (B,1,1)."""
  V = x.shape[-1]
  alpha_ts = alpha_t / alpha_s
  d_alpha = alpha_s - alpha_t
  oh = F.one_hot(xt, V)
  limiting = torch.ones(1, 1, V, dtype=x.dtype) / V
  x_at = torch.gather(x, -1, xt[..., None])
  num = (alpha_t * V * x * oh + (alpha_ts - alpha_t) * oh + d_alpha * x
         + (1 - alpha_ts) * (1 - alpha_s) * limiting)
  den = alpha_t * V * x_at + (1 - alpha_t)
  return num / den


def posterior_absorbing(x, mc_t, mc_s, mask_index):
  """MDLM posterior (diffusion.py:1399-1401)."""
  q = x * (mc_t - mc_s)
  q[:, :, mask_index] = mc_s[:, :, 0]
  return q / mc_t


def sample_categorical(probs, U, argmax_sampling=False, low_confidence=None):
  """`_sample_categorical` (diffusion.py:1738-1755) with injected U.
  `low_confidence` = eval.low_confidence_threshold when
  eval.low_confidence_sampling is on (:1739-1746), else None."""
  if low_confidence is not None:
    sorted_probs, sorted_idx = probs.sort(dim=-1, stable=True)
    mask_sorted = sorted_probs.cumsum(dim=-1) <= low_confidence
    mask = torch.zeros_like(mask_sorted)
    mask.scatter_(dim=-1, index=sorted_idx, src=mask_sorted)
    probs = probs.masked_fill(~mask, 0.0)
  if argmax_sampling:
    return probs.argmax(dim=-1)
  gumbel_norm = (1e-10 - (U + 1e-10).log()).to(probs.dtype)
  return (probs / gumbel_norm).argmax(dim=-1)


def _carry_over(xs, q_xs, xt, mask_index):
  copy_flag = xt != mask_index
  q_xs[copy_flag] = 0.0
  q_xs[copy_flag, xt[copy_flag]] = 1.0
  return torch.where(copy_flag, xt, xs), q_xs


# --------------------------------------------------------------------------
# per-step updates, from log-probs
# --------------------------------------------------------------------------
def ddpm_update(lp, xt, mc_t, mc_s, U, *, diffusion, mask_index,
                lp_dtype=None, argmax_sampling=False, use_float64=False,
                low_confidence=None):
  """`_ddpm_denoise` after the forward (diffusion.py:1393-1420)."""
  if use_float64:
    lp = lp.to(torch.float64)
  x = lp.exp() if use_float64 else _exp(lp, lp_dtype)   # float64: no bf16 rounding of exp
  if diffusion == 'absorbing_state':
    q = posterior_absorbing(x, mc_t, mc_s, mask_index)
  elif diffusion == 'uniform':
    q = posterior_uniform(x, xt, 1 - mc_s, 1 - mc_t)
  else:
    raise NotImplementedError(diffusion)
  xs = sample_categorical(q, U, argmax_sampling, low_confidence)
  if diffusion == 'absorbing_state':
    xs, q = _carry_over(xs, q, xt, mask_index)
  return xs, q


def cfg_update(lp_cond, lp_uncond, gamma, xt, mc_t, mc_s, U, *, diffusion,
               mask_index, lp_dtype=None, argmax_sampling=False,
               low_confidence=None):
  """`_cfg_denoise` after the forwards (diffusion.py:1456-1509).  Pass
  lp_cond=None for gamma==0 and lp_uncond=None for gamma==1."""
  if lp_cond is None or lp_uncond is None:
    lp = lp_uncond if lp_uncond is not None else lp_cond
    return ddpm_update(lp, xt, mc_t, mc_s, U, diffusion=diffusion,
                       mask_index=mask_index, lp_dtype=lp_dtype,
                       argmax_sampling=argmax_sampling,
                       low_confidence=low_confidence)
  if diffusion == 'absorbing_state':
    if lp_dtype is None:
      mix = gamma * lp_cond + (1 - gamma) * lp_uncond
      x = mix.softmax(dim=-1)
    else:  # every op on bf16 tensors rounds (SURVEY 8a' item 5)
      a = (gamma * lp_cond).to(lp_dtype).float()
      b = ((1 - gamma) * lp_uncond).to(lp_dtype).float()
      mix = (a + b).to(lp_dtype).float()
      x = mix.softmax(dim=-1).to(lp_dtype).float()
    q = posterior_absorbing(x, mc_t, mc_s, mask_index)
  elif diffusion == 'uniform':
    lq_u = posterior_uniform(_exp(lp_uncond, lp_dtype), xt,
                             1 - mc_s, 1 - mc_t).log()
    lq_c = posterior_uniform(_exp(lp_cond, lp_dtype), xt,
                             1 - mc_s, 1 - mc_t).log()
    q = (gamma * lq_c + (1 - gamma) * lq_u).softmax(dim=-1)
  else:
    raise NotImplementedError(diffusion)
  xs = sample_categorical(q, U, argmax_sampling, low_confidence)
  if diffusion == 'absorbing_state':
    xs, q = _carry_over(xs, q, xt, mask_index)
  return xs, q


def cbg_update(lp, clf_log_prob, gamma, xt, mc_t, mc_s, U, *, diffusion,
               mask_index, lp_dtype=None, argmax_sampling=False,
               low_confidence=None, clf_dtype=None):
  """`_cbg_denoise` after the forward and classifier term
  (diffusion.py:1587-1621).  clf_log_prob (B,L,V) fp32 values.  `clf_dtype`:
  dtype of the classifier term when it is not fp32 (then `gamma * term`, :1607,
  :1612, rounds to it).  The reference's term is fp32 on CPU AND on a GPU (exact:
  log_softmax under autocast(float32), classifier.py:325-327, measured on the B200;
  first-order: fp32 autograd result, :1540-1547), so this stays None."""
  gc = gamma * clf_log_prob
  if clf_dtype is not None:
    gc = gc.to(clf_dtype).float()
  if diffusion == 'absorbing_state':
    lq = lp + torch.log(1. - (mc_s / mc_t))
    lq[..., mask_index] = torch.log(mc_s / mc_t)[:, :, 0]
    g = gc + lq
    copy_flag = xt != mask_index
    g[copy_flag] = NEG_INF
    g[copy_flag, xt[copy_flag]] = 0.0
  elif diffusion == 'uniform':
    lq = posterior_uniform(_exp(lp, lp_dtype), xt, 1 - mc_s, 1 - mc_t).log()
    g = gc + lq
  else:
    raise NotImplementedError(diffusion)
  p = g.softmax(dim=-1)
  xs = sample_categorical(p, U, argmax_sampling, low_confidence)
  if diffusion == 'absorbing_state':
    xs = torch.where(xt != mask_index, xt, xs)
  return xs, p


def cbg_classifier_term_exact(clf_fn, xt, sigma, vocab_size,
                              conditioning_class, chunk=4096):
  """Exact D-CBG term (diffusion.py:1549-1585): log p(y | xt with position l
  replaced by v) for every (l, v).  `clf_fn(ids (n,L), sigma (n,)) ->
  (n, num_classes) log-probs`."""
  B, L = xt.shape
  out = torch.empty(B, L * vocab_size)
  for b in range(B):
    jumps = xt[b][None].repeat(L * vocab_size, 1)
    idx = torch.arange(L * vocab_size)
    jumps[idx, idx // vocab_size] = idx % vocab_size
    for s in range(0, jumps.shape[0], chunk):
      part = jumps[s:s + chunk]
      out[b, s:s + chunk] = clf_fn(
        part, sigma[b].expand(part.shape[0]))[:, conditioning_class]
  return out.view(B, L, vocab_size)


def cbg_classifier_term_approx(clf_fn_onehot, xt, sigma, vocab_size,
                               conditioning_class):
  """First-order D-CBG term (diffusion.py:1532-1548) via autograd on the CPU
  restatement.  `clf_fn_onehot(one_hot (B,L,V) f32, sigma) -> log-probs`."""
  oh = F.one_hot(xt, vocab_size).to(torch.float)
  with torch.enable_grad():
    oh.requires_grad_(True)
    lp = clf_fn_onehot(oh, sigma)
    lp[..., conditioning_class].sum().backward()
    grad = oh.grad
  ratio = (grad - (oh * grad).sum(dim=-1, keepdim=True)).detach()
  return (ratio + lp[..., conditioning_class][..., None, None]).detach()


# --------------------------------------------------------------------------
# AR
# --------------------------------------------------------------------------
def ar_select(next_log_probs, gumbel, low_confidence=None):
  """`_sample_token` (diffusion.py:1024-1038): Gumbel-max; with
  `eval.low_confidence_sampling` only the bottom `low_confidence` cumulative
  probability mass stays eligible (:1030-1037)."""
  if low_confidence is not None:
    probs = next_log_probs.softmax(dim=-1)
    sorted_probs, sorted_idx = probs.sort(dim=-1)
    mask_sorted = sorted_probs.cumsum(dim=-1) <= low_confidence
    mask = torch.zeros_like(mask_sorted)
    mask.scatter_(1, sorted_idx, mask_sorted)
    next_log_probs = next_log_probs.masked_fill(~mask, -float('inf'))
  return (next_log_probs + gumbel).argmax(-1)


def ar_cfg_mix(lp_cond, lp_uncond, gamma):
  """`_ar_cfg_denoise` tempered branch (diffusion.py:1233-1235)."""
  return (gamma * lp_cond + (1 - gamma) * lp_uncond).log_softmax(dim=-1)


# --------------------------------------------------------------------------
# whole sampler (denoiser = ref_dit restatement)
# --------------------------------------------------------------------------
class OracleSampler:
  """`Diffusion.sample` / `_diffusion_sample` (diffusion.py:1040-1083,
  1281-1376) with injected prior and noise.  `cfg` carries the reference's
  config keys as plain attributes."""

  def __init__(self, sd, *, vocab_size, mask_index, parameterization,
               diffusion, n_heads, time_conditioning, subs_masking=False,
               T=0, noise_type='loglinear', num_classes=None,
               gpu_rounding=False, argmax_sampling=False,
               clf_sd=None, clf_n_heads=None, use_float64=False,
               low_confidence=None):
    self.sd = sd
    self.V = vocab_size
    self.mask_index = mask_index
    self.parameterization = parameterization
    self.diffusion = diffusion
    self.n_heads = n_heads
    self.time_conditioning = time_conditioning
    self.subs_masking = subs_masking
    self.T = T
    self.noise_type = noise_type
    self.num_classes = num_classes
    self.argmax_sampling = argmax_sampling
    self.use_float64 = use_float64          # honoured by _ddpm_denoise only
    self.low_confidence = low_confidence    # threshold or None
    self.rounding = (ref_dit.Rounding.gpu_denoiser() if gpu_rounding
                     else ref_dit.Rounding.none())
    self.clf_rounding = (ref_dit.Rounding.gpu_classifier() if gpu_rounding
                         else ref_dit.Rounding.none())
    self.lp_dtype = torch.bfloat16 if gpu_rounding else None
    self.clf_sd = clf_sd
    self.clf_n_heads = clf_n_heads

  # Diffusion.forward (diffusion.py:350-365)
  def forward(self, x, sigma, cond=None):
    sigma = sigma.reshape(-1)
    if not self.time_conditioning:
      sigma = torch.zeros_like(sigma)
    logits = ref_dit.dit_forward(
      self.sd, x, sigma, cond, n_heads=self.n_heads,
      causal=self.parameterization == 'ar', rounding=self.rounding)
    return log_probs_from_logits(
      logits, x, parameterization=self.parameterization,
      mask_index=self.mask_index, subs_masking=self.subs_masking,
      lp_dtype=self.lp_dtype)

  def clf_log_probs(self, ids_or_oh, sigma):
    return ref_dit.classifier_log_probs(
      self.clf_sd, ids_or_oh, sigma, n_heads=self.clf_n_heads,
      time_conditioning=self.time_conditioning, rounding=self.clf_rounding)

  def step(self, xt, sigma_t, mc_t, mc_s, U, *, guidance=None, cond=None,
           gamma=1.0, condition=0, use_approx=False):
    """One reverse step -> (xs, q_xs, log-probs dict)."""
    kw = dict(diffusion=self.diffusion, mask_index=self.mask_index,
              lp_dtype=self.lp_dtype, argmax_sampling=self.argmax_sampling,
              low_confidence=self.low_confidence)
    if guidance is None:
      lp = self.forward(xt, sigma_t)
      xs, q = ddpm_update(lp, xt, mc_t, mc_s, U, use_float64=self.use_float64,
                          **kw)
      return xs, q, {'log_x_theta': lp}
    if guidance == 'cfg':
      lp_c = lp_u = None
      if gamma != 0.0:
        lp_c = self.forward(xt, sigma_t, cond)
      if gamma != 1.0:
        lp_u = self.forward(xt, sigma_t,
                            torch.full_like(cond, self.num_classes))
      xs, q = cfg_update(lp_c, lp_u, gamma, xt, mc_t, mc_s, U, **kw)
      return xs, q, {'log_x_theta_cond': lp_c, 'log_x_theta_uncond': lp_u}
    if guidance == 'cbg':
      lp = self.forward(xt, sigma_t)
      if use_approx:
        term = cbg_classifier_term_approx(
          self.clf_log_probs, xt, sigma_t, self.V, condition)
      else:
        term = cbg_classifier_term_exact(
          self.clf_log_probs, xt, sigma_t, self.V, condition)
      xs, q = cbg_update(lp, term, gamma, xt, mc_t, mc_s, U,
                         clf_dtype=None, **kw)   # the term is fp32 on a GPU too (ref_dit.classifier_log_probs)
      return xs, q, {'log_x_theta': lp, 'classifier_log_prob': term}
    raise NotImplementedError(
      f'Guidance method {guidance} not implemented.')

  def sample(self, xt, noise: typing.Callable[[int], torch.Tensor], steps,
             *, eps=1e-5, guidance=None, condition=0, gamma=1.0,
             use_approx=False, trace=None):
    """Full reverse process from prior `xt`; `noise(i)` returns step i's
    uniform tensor (B,L,V).  No cache: recomputing is result-identical
    (SURVEY.md section 7)."""
    B = xt.shape[0]
    sig, mct, mcs = step_scalars(steps, eps, self.T, self.noise_type)
    cond = None
    if guidance == 'cfg':
      cond = torch.full((B,), condition, dtype=torch.long)
    for i in range(steps):
      sigma_t = sig[i].expand(B)
      mc_t = mct[i].expand(B)[:, None, None]
      mc_s = mcs[i].expand(B)[:, None, None]
      xs, q, lps = self.step(xt, sigma_t, mc_t, mc_s, noise(i),
                             guidance=guidance, cond=cond, gamma=gamma,
                             condition=condition, use_approx=use_approx)
      if trace is not None:
        trace.append({'xt': xt, 'xs': xs, 'q_xs': q, **lps})
      xt = xs
    return xt


# --------------------------------------------------------------------------
# hierarchical exponential race (CPU restatement of csrc/sample_group.cu)
# --------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
  """Philox4x32-10 (Salmon et al., SC'11) on numpy uint32 arrays (broadcast);
  returns the four output words."""
  import numpy as np
  M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
  W0, W1 = 0x9E3779B9, 0xBB67AE85
  c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)))
  k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
  for _ in range(10):
    p0 = M0 * c0.astype(np.uint64)
    p1 = M1 * c2.astype(np.uint64)
    hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
    hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
    c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint32(k0), lo1, hi0 ^ c3 ^ np.uint32(k1), lo0
    k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
  return c0, c1, c2, c3


def group_race_select(probs, seed, step, group=8, threads=256):
  """Categorical draw by the hierarchical exponential race of the production
  sampler (csrc/sample_group.cu), restated on the CPU in float64 with the same
  Philox stream.  Distributionally identical to `sample_categorical`
  (diffusion.py:1748-1755): the minimum of independent exponentials with rates
  q_j over a group is Exp(sum q_j), and the arg-min inside the group is
  q_j / sum q_j independent of the minimum.

  group=8: groups of 8 consecutive columns.  group=64: the kernel's strided
  64-column groups (`group_race_select_wide`).

  probs: (rows, V) non-negative, any common scale per row.  Returns a dict:
  ids (rows,), group scores (rows, n_groups) in log2 units, and for the winning
  group the in-group cumulative sums / target so that a test can tell a real
  mismatch from a near-tie; `cols` (rows, group) = the columns of the winning
  group in pick order (-1 = past the vocabulary)."""
  import numpy as np
  if group == 64:
    return group_race_select_wide(probs, seed, step)
  p = probs.detach().double().cpu().numpy()
  rows, V = p.shape
  nG = (V + group - 1) // group
  pad = nG * group - V
  if pad:
    p = np.concatenate([p, np.zeros((rows, pad))], 1)
  p = p.reshape(rows, nG, group)
  S = p.sum(-1)
  G = np.arange(nG, dtype=np.int64)[None, :]
  r = np.arange(rows, dtype=np.uint64)[:, None]
  row_lo = (r & np.uint64(0xFFFFFFFF)).astype(np.uint32)
  row_hi = (r >> np.uint64(32)).astype(np.uint32)
  k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
  shift = int(np.log2(threads))
  ctr0 = ((G & (threads - 1)) | ((G >> (shift + 2)) << shift)).astype(np.uint32)
  words = np.stack(philox4x32_10(ctr0, row_lo, row_hi, np.uint32(step & 0xFFFFFFFF), k0, k1), -1)
  sel = ((G >> shift) & 3).astype(np.int64)
  w = np.take_along_axis(words, np.broadcast_to(sel[..., None], (rows, nG, 1)), -1)[..., 0]
  U = (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
  E = 1e-10 - np.log(U.astype(np.float64) + 1e-10)
  with np.errstate(divide='ignore'):
    score = np.log2(S) - np.log2(E)
  win = score.argmax(-1)
  pw = p[np.arange(rows), win]                               # (rows, group)
  w2 = philox4x32_10(win.astype(np.uint32), row_lo[:, 0],
                     row_hi[:, 0] | np.uint32(0x80000000), np.uint32(step & 0xFFFFFFFF), k0, k1)[0]
  u2 = (w2 >> np.uint32(8)).astype(np.float32).astype(np.float64) * 2.0 ** -24
  cum = np.cumsum(pw, -1)
  target = u2 * cum[:, -1]
  pick = (cum > target[:, None]).argmax(-1)
  cols = win[:, None] * group + np.arange(group)[None, :]
  cols = np.where(cols < V, cols, -1)
  return dict(ids=torch.from_numpy(win * group + pick), score=score, win=win, cum=cum,
              target=target, u2=u2, cols=cols)


def wide_group_columns(V, threads=128):
  """Column map of the kernel's 64-column groups (race_group = 64): with vector index
  v = j // 8, slot t = v % threads, kk = v // threads, sweep k8 = kk // 8, group
  G = k8 * threads + t holds columns 8 (t + threads (8 k8 + s)) + e in pick order
  (s, e).  Returns (n_groups, 64) int64 with -1 past the vocabulary."""
  import numpy as np
  n_vec = (V + 7) // 8
  n_k8 = (n_vec + 8 * threads - 1) // (8 * threads)
  G = np.arange(n_k8 * threads, dtype=np.int64)
  t, k8 = G % threads, G // threads
  s = np.arange(8, dtype=np.int64)
  e = np.arange(8, dtype=np.int64)
  v = t[:, None] + threads * (8 * k8[:, None] + s[None, :])           # (nG, 8)
  cols = (8 * v[:, :, None] + e[None, None, :]).reshape(len(G), 64)
  return np.where(cols < V, cols, -1)


def group_race_select_wide(probs, seed, step, threads=128):
  """`group_race_select` for the strided 64-column groups of race_group = 64
  (csrc/sample_group.cu, WIDE): group exponential from Philox counter
  (t | ((k8 >> 2) << 8), row_lo, row_hi | 0x40000000, step), word k8 & 3; in-group
  pick from counter (G, row_lo, row_hi | 0x80000000, step), word 0."""
  import numpy as np
  p = probs.detach().double().cpu().numpy()
  rows, V = p.shape
  cols = wide_group_columns(V, threads)                              # (nG, 64)
  nG = cols.shape[0]
  pg = np.where(cols[None] >= 0, p[:, np.maximum(cols, 0)], 0.0)     # (rows, nG, 64)
  S = pg.sum(-1)
  G = np.arange(nG, dtype=np.int64)[None, :]
  t, k8 = G % threads, G // threads
  r = np.arange(rows, dtype=np.uint64)[:, None]
  row_lo = (r & np.uint64(0xFFFFFFFF)).astype(np.uint32)
  row_hi = (r >> np.uint64(32)).astype(np.uint32)
  k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
  ctr0 = (t | ((k8 >> 2) << 8)).astype(np.uint32)
  words = np.stack(philox4x32_10(ctr0, row_lo, row_hi | np.uint32(0x40000000),
                                 np.uint32(step & 0xFFFFFFFF), k0, k1), -1)
  sel = (k8 & 3).astype(np.int64)
  w = np.take_along_axis(words, np.broadcast_to(sel[..., None], (rows, nG, 1)), -1)[..., 0]
  U = (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
  E = 1e-10 - np.log(U.astype(np.float64) + 1e-10)
  with np.errstate(divide='ignore'):
    score = np.log2(S) - np.log2(E)
  win = score.argmax(-1)
  pw = pg[np.arange(rows), win]                                      # (rows, 64)
  w2 = philox4x32_10(win.astype(np.uint32), row_lo[:, 0],
                     row_hi[:, 0] | np.uint32(0x80000000), np.uint32(step & 0xFFFFFFFF), k0, k1)[0]
  u2 = (w2 >> np.uint32(8)).astype(np.float32).astype(np.float64) * 2.0 ** -24
  cum = np.cumsum(pw, -1)
  target = u2 * cum[:, -1]
  pick = (cum > target[:, None]).argmax(-1)
  wcols = cols[win]                                                  # (rows, 64)
  ids = wcols[np.arange(rows), pick]
  return dict(ids=torch.from_numpy(ids), score=score, win=win, cum=cum, target=target,
              u2=u2, cols=wcols)
