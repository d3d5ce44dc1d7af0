"""Reverse-diffusion sampler: host-side mirror of the reference `Diffusion`
sampling surface (`diffusion.py:1040-1083, 1281-1624, 1738-1755`).

Same constructor arguments, config keys, method names, argument meanings and
error behaviour as the reference for the sampling path:
`Diffusion(config, tokenizer)`, `.sample(eps)`, `._diffusion_sample(...)`,
`._ddpm_denoise / ._cfg_denoise / ._cbg_denoise(...) -> (xs, q_xs, cache)`,
`.forward(x, sigma, cond)`.  Underneath, each step is a handful of launches of
the sm_100a kernels behind the C ABI: one DiT forward (the CFG cond/uncond pair
batched into a single 2B-row forward), then ONE fused kernel that turns the
logits into the next token ids.  `sample()` captures one step in a CUDA graph
and replays it `sampling.steps` times with all per-step scalars read from
device tables.
"""
from __future__ import annotations

import collections
import gc
import typing

import torch

from . import _lib, config as config_lib, noise_schedule, weights
from ._lib import check, lib, ptr, stream_ptr
from .dit import DIT, ModSpec


_MAX_STEP_GRAPHS = 2


def _capture(body):
  """Captures `body()` into a new CUDA graph.  Dead Python cycles that hold an older
  `torch.cuda.CUDAGraph` (a discarded model and its step graphs) must not be collected
  WHILE the stream is capturing -- destroying a graph is not a capturable operation and
  invalidates the capture -- so they are collected first and the cyclic collector stays
  off until the capture ends (torch.cuda.graph no longer collects on entry)."""
  gc.collect()
  was_enabled = gc.isenabled()
  gc.disable()
  try:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode='thread_local'):
      body()
  finally:
    if was_enabled:
      gc.enable()
  return g


def _attach_stats(logits, stats, row_offset):
  """The vocab GEMM's log-softmax partials travel WITH the logits tensor they
  describe (as an attribute, so the cache dicts of the per-step API carry them and
  their lifetime is the tensor's), tagged with the tensor's version counter: an
  in-place write to the logits afterwards invalidates them."""
  logits._ddg_stats = (stats, int(row_offset), logits._version)
  return logits


def _get_stats(logits):
  tag = getattr(logits, '_ddg_stats', None)
  if tag is None or tag[2] != logits._version:
    return None
  return tag[0], tag[1]


class Diffusion:

  def __init__(self, config, tokenizer, device=None):
    _lib.require_device()
    self.config = config
    self.tokenizer = tokenizer
    self.device = torch.device(device if device is not None else 'cuda')
    # diffusion.py:80-95
    self.vocab_size = tokenizer.vocab_size
    if (not hasattr(tokenizer, 'mask_token') or tokenizer.mask_token is None):
      self.mask_index = self.vocab_size
      self.vocab_size += 1
    else:
      self.mask_index = tokenizer.mask_token_id
    self.parameterization = config.parameterization
    self.diffusion = config.diffusion
    if self.parameterization != 'ar' and self.diffusion not in (
        'absorbing_state', 'uniform'):
      raise NotImplementedError(
        f'Diffusion type {self.diffusion} not implemented.')
    self.T = config.T
    self.subs_masking = config.subs_masking
    self.time_conditioning = config.time_conditioning
    config_lib.validate(config)
    if config.backbone != 'dit':
      raise NotImplementedError(
        f'Backbone {config.backbone} not implemented.')
    self.backbone = DIT(config, vocab_size=self.vocab_size, device=self.device)
    self.noise = noise_schedule.get_noise(config)
    self.neg_infinity = -1_000_000.0
    self.classifier_model = None
    self._classifier_path = None
    # knobs that are not reference config keys
    self.round_bf16 = True      # reproduce the reference's bf16 log-softmax/exp
    self.fast_math = True       # ex2/lg2 intrinsics in the fused sampler
    self.ar_kv_cache = True     # AR sampler: incremental decode against a KV
                                # cache; False = re-encode the prefix per
                                # token like the reference (diffusion.py:1132)
    self.race_group = 64        # Philox path: hierarchical race over 64-column
                                # groups (same distribution, 64x less noise
                                # work); 8 = 8-column groups; 0 = one
                                # exponential per vocab entry
    self.logits_cache = True    # MDLM without time conditioning: a sequence whose tokens did
                                # not change keeps its logits; only changed sequences go
                                # through the network (the per-sequence form of the
                                # reference's sampling.use_cache, diffusion.py:1371-1374;
                                # ids identical to recomputing everything)
    self.return_q_xs = False    # materialise q_xs (B, L, V) fp32 per step
    self.use_cuda_graph = True
    self.noise_fn = None        # i -> injected U (B, L, V); None -> Philox
    self.prior_fn = None        # (B, L) -> injected prior ids
    self.seed = int(config.get('seed', 1))
    self._call_count = 0
    self._loop = None
    self._step_index = 0        # Philox counter word of the next eager sampler launch
    self._step_graphs = collections.OrderedDict()   # LRU, at most _MAX_STEP_GRAPHS

  # ---------------------------------------------------------------- weights
  def load_state_dict(self, state_dict, strict=True):
    self.backbone.load_state_dict(state_dict, prefix='backbone.')
    return self

  @classmethod
  def load_from_checkpoint(cls, checkpoint_path, tokenizer, config, **kw):
    """`Diffusion.load_from_checkpoint` (main.py:39-47): EMA shadow params
    are used unless `eval.disable_ema` (diffusion.py:1054-1055)."""
    self = cls(config, tokenizer, **kw)
    sd = weights.load_checkpoint_state(
      checkpoint_path, use_ema=not config.eval.disable_ema)
    return self.load_state_dict(sd)

  def set_classifier(self, classifier_model):
    self.classifier_model = classifier_model

  def to(self, device):
    return self

  def eval(self):
    return self

  # ---------------------------------------------------------------- helpers
  def _process_sigma(self, sigma):
    # diffusion.py:339-348
    if sigma is None:
      assert self.parameterization == 'ar'
      return sigma
    if sigma.ndim > 1:
      sigma = sigma.squeeze(-1)
    if not self.time_conditioning:   # one shared zero (stride 0): a single conditioning row
      sigma = torch.zeros(1, dtype=sigma.dtype, device=sigma.device).expand(sigma.shape[0])
    assert sigma.ndim == 1, sigma.shape
    return sigma

  @property
  def _mask_logit(self):
    return self.parameterization == 'subs' or bool(self.subs_masking)

  def _diffusion_code(self):
    if self.parameterization == 'ar':
      return _lib.AR
    if self.diffusion == 'absorbing_state':
      return _lib.ABSORBING
    if self.diffusion == 'uniform':
      return _lib.UNIFORM
    raise NotImplementedError(
      f'Diffusion type {self.diffusion} not implemented.')

  def _select_code(self):
    ev = self.config.get('eval') or {}
    return (_lib.SELECT_ARGMAX if ev.get('argmax_sampling', False)
            else _lib.SELECT_RACE)

  @property
  def _low_confidence(self):
    """eval.low_confidence_sampling (diffusion.py:1738-1746)."""
    ev = self.config.get('eval') or {}
    return bool(ev.get('low_confidence_sampling', False))

  @property
  def _use_float64(self):
    """sampling.use_float64: honoured, like the reference, only by
    `_ddpm_denoise` and the AR sampler (diffusion.py:1393-1394, 1134-1135)."""
    return bool(self.config.sampling.get('use_float64', False))

  def _q_buffer(self, B, L):
    """(B, L, V) fp32 posterior: returned when `return_q_xs`, and the input of
    the low-confidence draw."""
    if self.return_q_xs or self._low_confidence:
      return torch.empty(B, L, self.vocab_size, dtype=torch.float32,
                         device=self.device)
    return None

  def _low_confidence_draw(self, q, xt, xs, noise, carry):
    """Replaces `xs` by the draw restricted to the low-confidence set
    (diffusion.py:1738-1755), then the absorbing carry-over."""
    B, L, V = q.shape
    thr = float(self.config.eval.get('low_confidence_threshold', 0.3))
    check(lib.ddg_low_confidence_select(
      ptr(q), V, B * L, V, thr, ptr(noise), V if noise is not None else 0,
      self._select_code(), self.seed + 7919 * self._call_count, 0,
      int(self._step_index), 0, ptr(xt) if carry else 0,
      self.mask_index if carry else -1, ptr(xs), stream_ptr()))

  @property
  def _stats_mask_col(self):
    return self.mask_index if self._mask_logit else -1

  def _logits(self, x, sigma, cond=None):
    """Backbone call of Diffusion.forward (diffusion.py:350-354): raw bf16
    logits, shape (B, L, V) (strided view).  The vocab GEMM also leaves its
    log-softmax partials for the fused sampler."""
    sigma = self._process_sigma(sigma)
    B, L = x.shape
    stats = self.backbone.alloc_stats(B * L)
    logits = self.backbone.forward(x, sigma, cond, stats_out=stats,
                                   stats_mask_col=self._stats_mask_col)
    return _attach_stats(logits, stats, 0)

  def forward(self, x, sigma, cond=None, x_emb=None, **kwargs):
    """Returns log-probs (B, L, V) fp32 like `Diffusion.forward`
    (diffusion.py:350-365).  The sampling path never calls this: the fused
    kernel consumes the raw logits."""
    if x_emb is not None or kwargs:
      raise NotImplementedError('x_emb / inference_params are not on the DiT sampling path')
    logits = self._logits(x, sigma, cond)
    B, L, V = logits.shape
    ids = x.to(self.device, torch.int64).contiguous()
    out = torch.empty(B, L, V, dtype=torch.float32, device=self.device)
    check(lib.ddg_log_probs(
      logits.data_ptr(), logits.stride(1), _lib.BF16, ptr(ids), B * L, V,
      1 if self.parameterization == 'subs' else 0, 1 if self._mask_logit else 0,
      self.mask_index, 1 if self.round_bf16 else 0, ptr(out), V, stream_ptr()))
    return out

  def _sample_prior(self, *batch_dims):
    """diffusion.py:1006-1013 (uniform: in-kernel Philox randint)."""
    if self.prior_fn is not None:
      return self.prior_fn(*batch_dims).to(self.device, torch.int64).contiguous()
    xt = torch.empty(*batch_dims, dtype=torch.int64, device=self.device)
    check(lib.ddg_sample_prior(ptr(xt), xt.numel(), self._diffusion_code(),
                               self.vocab_size, self.mask_index,
                               self.seed + 7919 * self._call_count,
                               stream_ptr()))
    return xt

  def _step_scalars(self, steps, eps):
    """(sigma_t, move_chance_t, move_chance_s) per step, formed with the same
    fp32 torch ops as `_diffusion_sample` (diffusion.py:1293-1323)."""
    timesteps = torch.linspace(1, eps, steps + 1)
    dt = (1 - eps) / steps
    t = timesteps[:-1].clone()
    if self.T > 0:
      t = (t * self.T).to(torch.int)
      t = t / self.T
      t += (1 / self.T)
    sigma_t = self.noise.total_noise(t)
    sigma_s = self.noise.total_noise(t - dt)
    return sigma_t, 1 - torch.exp(-sigma_t), 1 - torch.exp(-sigma_s)

  @staticmethod
  def _stats_of(a, b=None):
    """Log-softmax partials the vocab GEMM left for these logits tensors, if they
    travel with them (`_attach_stats`) and the tensors were not written since."""
    sa = _get_stats(a)
    if sa is None:
      return None
    if b is None:
      return (sa[0], sa[1], 0)
    sb = _get_stats(b)
    if sb is None or sb[0] is not sa[0]:
      return None
    return (sa[0], sa[1], sb[1])

  def _launch_sampler(self, *, logits_a, ld_a, logits_b=None, ld_b=0, clf=None,
                      noise=None, xt, xs, mc_t=None, mc_s=None, rows, seq_len,
                      mc_seq_stride=0, mc_step_stride=0, step_ptr=None, step=0,
                      guidance=_lib.GUIDE_NONE, gamma=1.0, q_out=None,
                      select=None, diffusion=None, logits_dtype=_lib.BF16,
                      seed_ptr=None, stats=None, compute_f64=False,
                      clf_round_bf16=False, seq_changed=None):
    V = self.vocab_size
    # log-softmax partials of these logits, if our own vocab GEMM made them:
    # stats = (tensor (chunks, rows_total, 2), row offset of a, row offset of b); they
    # are handed in by the caller together with the logits (never looked up by address)
    st_a = st_b = st_ld = st_chunks = 0
    if stats is not None:
      st, off_a, off_b = stats
      st_a, st_b = st.data_ptr() + off_a * 8, st.data_ptr() + off_b * 8
      st_ld, st_chunks = st.shape[1], st.shape[0]
    a = _lib.SampleStepArgs(
      logits_a=logits_a, logits_b=logits_b or 0, clf=ptr(clf), noise=ptr(noise),
      xt=ptr(xt), mc_t=ptr(mc_t), mc_s=ptr(mc_s), step_ptr=ptr(step_ptr),
      seed_ptr=ptr(seed_ptr), xs=ptr(xs), q_out=ptr(q_out), rows=rows, ld_a=ld_a, ld_b=ld_b,
      ld_clf=clf.stride(-2) if clf is not None else 0,
      ld_noise=(noise.stride(-2) if noise is not None else 0),
      ld_q=V if q_out is not None else 0, seq_len=seq_len, vocab=V,
      mc_seq_stride=mc_seq_stride, mc_step_stride=mc_step_stride,
      logits_dtype=logits_dtype,
      diffusion=self._diffusion_code() if diffusion is None else diffusion,
      guidance=guidance, subs=1 if self.parameterization == 'subs' else 0,
      mask_logit=1 if self._mask_logit else 0, mask_index=self.mask_index,
      round_bf16=1 if self.round_bf16 else 0,
      select=self._select_code() if select is None else select,
      fast_math=1 if (self.fast_math and noise is None) else 0,
      gamma=float(gamma), seed=self.seed + 7919 * self._call_count,
      step=int(step), stats_chunks=st_chunks, stats_a=st_a,
      stats_b=st_b if logits_b else 0, stats_ld=st_ld,
      race_group=int(self.race_group), compute_f64=1 if compute_f64 else 0,
      clf_round_bf16=1 if clf_round_bf16 else 0, seq_changed=ptr(seq_changed))
    check(lib.ddg_sample_step(a, stream_ptr()))

  def _mc(self, mc):
    return mc.to(self.device, torch.float32).reshape(-1).contiguous()

  def _step_graph_ok(self, cache):
    """Per-step API fast path: the whole step (forward + fused sampler) is one
    captured CUDA graph fed through static input buffers."""
    return (self.use_cuda_graph and cache is None and self.noise_fn is None
            and not self.return_q_xs and not self._low_confidence
            and not self._use_float64)

  def _get_step_graph(self, kind, B, L, gamma):
    key = (kind, B, L, float(gamma), self.round_bf16, self.fast_math,
           self.race_group)
    sg = self._step_graphs.get(key)
    if sg is None:
      # each graph owns its logits buffer (8 GB at B=512, L=128, V=30522 for cfg2): keep
      # the most recently used few, so a gamma sweep does not grow HBM use without bound
      while len(self._step_graphs) >= _MAX_STEP_GRAPHS:
        self._step_graphs.popitem(last=False)
      sg = self._step_graphs[key] = _StepGraph(self, kind, B, L, float(gamma))
    else:
      self._step_graphs.move_to_end(key)
    return sg

  def _step_noise(self, B, L):
    if self.noise_fn is None:
      return None
    U = self.noise_fn(self._step_index).to(self.device, torch.float32)
    return U.reshape(B * L, self.vocab_size).contiguous()

  # ---------------------------------------------------------------- per step
  def _ddpm_denoise(self, xt, time_conditioning, move_chance_t, move_chance_s,
                    cache=None):
    """diffusion.py:1378-1420 -> (xs, q_xs | None, cache)."""
    B, L = xt.shape
    xt = xt.to(self.device, torch.int64).contiguous()
    if self._step_graph_ok(cache):
      sg = self._get_step_graph('ddpm', B, L, 1.0)
      xs = sg.run(xt, self._process_sigma(time_conditioning), None,
                  move_chance_t, move_chance_s)
      return xs, None, {'log_x_theta': sg.logits_view(0)}
    logits = (cache['log_x_theta'] if cache is not None
              else self._logits(xt, time_conditioning, None))
    xs = torch.empty_like(xt)
    q = self._q_buffer(B, L)
    mct, mcs = self._mc(move_chance_t), self._mc(move_chance_s)
    noise = self._step_noise(B, L)
    self._launch_sampler(
      logits_a=logits.data_ptr(), ld_a=logits.stride(1), xt=xt, xs=xs, mc_t=mct,
      mc_s=mcs, rows=B * L, seq_len=L, mc_seq_stride=1 if mct.numel() > 1 else 0,
      noise=noise, q_out=q, step=self._step_index, stats=self._stats_of(logits),
      compute_f64=self._use_float64)
    if self._low_confidence:
      self._low_confidence_draw(q, xt, xs, noise,
                                carry=self.diffusion == 'absorbing_state')
    self._step_index += 1       # successive API calls draw fresh noise
    return xs, q if self.return_q_xs else None, {'log_x_theta': logits}

  def _cfg_denoise(self, cond, gamma, xt, time_conditioning, move_chance_t,
                   move_chance_s, cache=None):
    """diffusion.py:1422-1509 -> (xs, q_xs | None, cache).  gamma == 0 / 1 run
    one forward; otherwise cond and uncond are one batched 2B-row forward."""
    B, L = xt.shape
    xt = xt.to(self.device, torch.int64).contiguous()
    cond = cond.to(self.device, torch.int64)
    mask_cond = torch.ones_like(cond) * self.config.data.num_classes
    lc = lu = None
    if self._step_graph_ok(cache):
      sigma = self._process_sigma(time_conditioning)
      if gamma == 0.0 or gamma == 1.0:
        sg = self._get_step_graph('cfg1', B, L, gamma)
        xs = sg.run(xt, sigma, mask_cond if gamma == 0.0 else cond,
                    move_chance_t, move_chance_s)
        one = sg.logits_view(0)
        return xs, None, {'log_x_theta_uncond': one if gamma == 0.0 else None,
                          'log_x_theta_cond': one if gamma == 1.0 else None}
      sg = self._get_step_graph('cfg2', B, L, gamma)
      xs = sg.run(xt, sigma, torch.cat([cond, mask_cond]), move_chance_t,
                  move_chance_s)
      return xs, None, {'log_x_theta_uncond': sg.logits_view(1),
                        'log_x_theta_cond': sg.logits_view(0)}
    if cache is not None:
      lu, lc = cache['log_x_theta_uncond'], cache['log_x_theta_cond']
    elif gamma == 0.0:
      lu = self._logits(xt, time_conditioning, mask_cond)
    elif gamma == 1.0:
      lc = self._logits(xt, time_conditioning, cond)
    else:
      sigma = self._process_sigma(time_conditioning).to(self.device, torch.float32)
      mod = ModSpec(n_mod=2 * B, seqs_per_mod=1, sigma=torch.cat([sigma, sigma]),
                    sigma_row_stride=1, labels=torch.cat([cond, mask_cond]))
      buf = self.backbone.alloc_logits(2 * B * L)
      stats = self.backbone.alloc_stats(2 * B * L)
      self.backbone.run(xt, 2 * B, L, mod, buf, ids_period=B * L,
                        stats_out=stats, stats_mask_col=self._stats_mask_col)
      both = buf.view(2, B, L, self.backbone.vocab_ld)[..., :self.vocab_size]
      lc = _attach_stats(both[0], stats, 0)
      lu = _attach_stats(both[1], stats, B * L)
    xs = torch.empty_like(xt)
    q = self._q_buffer(B, L)
    mct, mcs = self._mc(move_chance_t), self._mc(move_chance_s)
    noise = self._step_noise(B, L)
    common = dict(xt=xt, xs=xs, mc_t=mct, mc_s=mcs, rows=B * L, seq_len=L,
                  mc_seq_stride=1 if mct.numel() > 1 else 0,
                  noise=noise, q_out=q, step=self._step_index)
    if lc is None or lu is None:
      one = lu if lu is not None else lc
      self._launch_sampler(logits_a=one.data_ptr(), ld_a=one.stride(1),
                           stats=self._stats_of(one), **common)
    else:
      self._launch_sampler(logits_a=lc.data_ptr(), ld_a=lc.stride(1),
                           logits_b=lu.data_ptr(), ld_b=lu.stride(1),
                           stats=self._stats_of(lc, lu),
                           guidance=_lib.GUIDE_CFG, gamma=gamma, **common)
    if self._low_confidence:
      self._low_confidence_draw(q, xt, xs, noise,
                                carry=self.diffusion == 'absorbing_state')
    self._step_index += 1
    return xs, q if self.return_q_xs else None, {
      'log_x_theta_uncond': lu, 'log_x_theta_cond': lc}

  def _classifier_term(self, classifier_model, xt, time_conditioning,
                       conditioning_class, use_approx, raw_ok=False, out=None, live=None):
    """(B, L, V) fp32 classifier log-prob term of D-CBG.  `raw_ok`: the caller only
    samples from it (the term is not handed back), so the row-constant part of the
    first-order term, which cancels in the guided softmax, may be left out.
    `out` / `live` (the captured step loop): preallocated flat term buffer, and a device
    int32 scalar that keeps the buffer's previous contents when it reads 0."""
    if use_approx:
      # first order (diffusion.py:1532-1548): one classifier forward + backward
      term, _ = classifier_model.first_order_term(xt, time_conditioning,
                                                  conditioning_class, finalize=not raw_ok,
                                                  out=out, live=live)
      return term
    # exact: every single-token edit of xt (diffusion.py:1549-1585), generated
    # chunk by chunk on the device and never materialised as a whole.
    B, L = xt.shape
    V = self.vocab_size
    total = B * L * V
    term = (out.view(-1)[:total] if out is not None
            else torch.empty(total, dtype=torch.float32, device=self.device))
    sigma = self._process_sigma(time_conditioning).to(self.device, torch.float32)
    chunk = max(L * V, (classifier_model.chunk_sequences // (L * V)) * (L * V))
    for c0 in range(0, total, chunk):
      n = min(chunk, total - c0)
      ids = torch.empty(n, L, dtype=torch.int64, device=self.device)
      check(lib.ddg_cbg_candidates(ptr(xt), L, V, c0, n, ptr(ids), stream_ptr()))
      # whole sequences per chunk: the L * V edits of a sequence share its sigma, so the
      # classifier gets one conditioning row per sequence (or one in all when every
      # sequence shares t), not one per candidate (diffusion.py:1584 repeats sigma instead)
      b0, b1 = c0 // (L * V), (c0 + n) // (L * V)
      lp = classifier_model.get_log_probs(ids, sigma[b0:b1], seqs_per_mod=L * V)
      new = lp[:, conditioning_class]
      if live is not None:
        new = torch.where(live > 0, new, term[c0:c0 + n])
      term[c0:c0 + n] = new
    return term.view(B, L, V)

  def _cbg_denoise(self, conditioning_class, gamma, classifier_model, xt,
                   time_conditioning, move_chance_t, move_chance_s,
                   use_approx=False, cache=None, _sampling_only=False):
    """diffusion.py:1511-1624 -> (xs, guided_probs | None, cache).
    `_sampling_only` (the `sample()` loop, which drops the cache): see
    `_classifier_term(raw_ok=...)`; the classifier term in the returned cache then
    lacks its row-constant part (good for re-use by this loop only)."""
    B, L = xt.shape
    xt = xt.to(self.device, torch.int64).contiguous()
    if cache is not None:
      logits, term = cache['log_x_theta'], cache['classifier_log_prob']
    else:
      logits = self._logits(xt, time_conditioning, None)
      term = self._classifier_term(classifier_model, xt, time_conditioning,
                                   conditioning_class, use_approx,
                                   raw_ok=_sampling_only and self.noise_fn is None)
    xs = torch.empty_like(xt)
    q = self._q_buffer(B, L)
    mct, mcs = self._mc(move_chance_t), self._mc(move_chance_s)
    noise = self._step_noise(B, L)
    self._launch_sampler(
      logits_a=logits.data_ptr(), ld_a=logits.stride(1),
      clf=term if term.stride(-1) == 1 else term.contiguous(),
      xt=xt, xs=xs, mc_t=mct, mc_s=mcs, rows=B * L, seq_len=L,
      mc_seq_stride=1 if mct.numel() > 1 else 0, noise=noise,
      q_out=q, guidance=_lib.GUIDE_CBG, gamma=gamma, step=self._step_index,
      stats=self._stats_of(logits),
      # the classifier term is fp32 in the reference on a GPU too (its log_softmax runs under
      # autocast(float32), classifier.py:325-327; measured: tests/golden/*_cbg_gpu.npz), so
      # gamma * term is an fp32 product: no bf16 rounding of it
      clf_round_bf16=False)
    if self._low_confidence:
      self._low_confidence_draw(q, xt, xs, noise,
                                carry=self.diffusion == 'absorbing_state')
    self._step_index += 1
    return xs, q if self.return_q_xs else None, {
      'log_x_theta': logits,
      'classifier_log_prob': term}

  # ---------------------------------------------------------------- loops
  @torch.no_grad()
  def _diffusion_sample(self, classifier_model=None, cond=None, eps=1e-5):
    """diffusion.py:1281-1376.  Philox noise + no diagnostics -> one captured
    step replayed `steps` times; injected noise -> eager per-step calls."""
    B = self.config.sampling.batch_size
    L = self.config.model.length
    steps = self.config.sampling.steps
    guidance = self.config.get('guidance')
    method = None if guidance is None else guidance.method
    if method not in (None, 'cfg', 'cbg'):
      raise NotImplementedError(f'Guidance method {method} not implemented.')
    xt = self._sample_prior(B, L)
    sigma_t, mc_t, mc_s = self._step_scalars(steps, eps)
    mixed_cond = cond is not None and not bool((cond == cond.reshape(-1)[0]).all())
    if self.noise_fn is not None or self.return_q_xs \
        or not self.use_cuda_graph or self._low_confidence or mixed_cond \
        or (self._use_float64 and method is None):
      return self._eager_loop(xt, sigma_t, mc_t, mc_s, method, guidance,
                              classifier_model, cond)
    if method == 'cbg':
      self.classifier_model = classifier_model
    loop = self._get_loop(B, L, method, guidance)
    return loop.run(xt, sigma_t, mc_t, mc_s, cond)

  def _eager_loop(self, xt, sigma_t, mc_t, mc_s, method, guidance,
                  classifier_model, cond):
    """Step loop of `_diffusion_sample` through the per-step API, with the reference's
    cache rule (diffusion.py:1371-1374): with `sampling.use_cache` the (logits,
    classifier term) of a step are reused by the next one while NO token changed."""
    B = xt.shape[0]
    dev = self.device
    use_cache = bool(self.config.sampling.get('use_cache', False))
    cache = None
    for i in range(sigma_t.numel()):
      self._step_index = i
      sig = sigma_t[i].to(dev).expand(B)
      mct = mc_t[i].to(dev).expand(B)[:, None, None]
      mcs = mc_s[i].to(dev).expand(B)[:, None, None]
      if method is None:
        xs, _, new_cache = self._ddpm_denoise(xt, sig, mct, mcs, cache=cache)
      elif method == 'cfg':
        xs, _, new_cache = self._cfg_denoise(cond, guidance.gamma, xt, sig, mct, mcs,
                                             cache=cache)
      else:
        xs, _, new_cache = self._cbg_denoise(
          guidance.condition, guidance.gamma, classifier_model, xt, sig, mct,
          mcs, use_approx=guidance.get('use_approx', False), cache=cache,
          _sampling_only=True)
      cache = new_cache if (use_cache and torch.equal(xs, xt)) else None
      xt = xs
    self._step_index = 0
    return xt

  def _get_loop(self, B, L, method, guidance):
    gamma = None if guidance is None else float(guidance.gamma)
    key = (B, L, method, gamma, self.round_bf16, self.fast_math,
           self.race_group, self.logits_cache)
    if method == 'cbg':
      key += (int(guidance.condition), bool(guidance.get('use_approx', False)),
              id(self.classifier_model))
    if self._loop is None or self._loop.key != key:
      self._loop = _GraphLoop(self, B, L, method, gamma, key)
    return self._loop

  def sample(self, eps=1e-5):
    """`Diffusion.sample` (diffusion.py:1040-1083) -> (batch, length) int64."""
    guidance = self.config.get('guidance')
    classifier_model, cond = None, None
    if guidance is not None:
      if guidance.method == 'cfg':
        cond = (torch.ones(self.config.sampling.batch_size, device=self.device)
                * guidance.condition).to(torch.long)
      if guidance.method in ('cbg', 'nos', 'fudge', 'pplm'):
        classifier_model = self._get_classifier(guidance)
    self._call_count += 1
    if self.parameterization == 'ar':
      return self._ar_sample(classifier_model=classifier_model, cond=cond)
    return self._diffusion_sample(classifier_model=classifier_model, cond=cond,
                                  eps=eps)

  def _get_classifier(self, guidance):
    if self.classifier_model is not None:
      return self.classifier_model
    from .classifier import Classifier   # loaded once, not per sample() call
    path = guidance.classifier_checkpoint_path
    self.classifier_model = Classifier.load_from_checkpoint(
      path, tokenizer=self.tokenizer, config=self.config, device=self.device)
    return self.classifier_model

  @torch.no_grad()
  def _ar_sample(self, classifier_model=None, cond=None):
    """diffusion.py:1085-1186 (+ _ar_cfg_denoise :1188-1236) for the DiT
    backbone.  The reference re-encodes the growing prefix for every token
    (diffusion.py:1132); here each step pushes ONE token per sequence through the
    network against a KV cache (`ar_kv_cache`, same logits up to summation
    order), or -- `ar_kv_cache = False` -- runs a fixed full-length causal
    forward whose row i is the prefix re-encode."""
    guidance = self.config.get('guidance')
    method = None if guidance is None else guidance.method
    if method not in (None, 'cfg'):
      raise NotImplementedError(f'Guidance method {method} not implemented.')
    B = self.config.sampling.batch_size
    L = self.config.model.length
    V, ld = self.vocab_size, self.backbone.vocab_ld
    x = torch.zeros(B, L, dtype=torch.int64, device=self.device)
    x[:, 0] = self.tokenizer.bos_token_id
    two = method == 'cfg' and guidance.gamma not in (0.0, 1.0)
    n_fwd = 2 if two else 1
    buf = None if self.ar_kv_cache else self.backbone.alloc_logits(n_fwd * B * L)
    y = torch.empty(B, dtype=torch.int64, device=self.device)
    # eval.low_confidence_sampling in `_sample_token` (diffusion.py:1030-1037): the fused
    # kernel leaves softmax(next_log_probs), the low-confidence kernel draws from its kept set
    q_next = (torch.empty(B, V, dtype=torch.float32, device=self.device)
              if self._low_confidence else None)
    labels = None
    if method == 'cfg':
      mask_cond = torch.ones_like(cond) * self.config.data.num_classes
      labels = (torch.cat([cond, mask_cond]) if two
                else (mask_cond if guidance.gamma == 0.0 else cond))
    mod = ModSpec(n_mod=n_fwd * B, seqs_per_mod=1, labels=labels)
    kv = step_logits = None
    if self.ar_kv_cache:
      kv = self.backbone.alloc_kv_cache(n_fwd * B, L)
      step_logits = self.backbone.alloc_logits(n_fwd * B)
    for i in range(L - 1):
      if kv is not None:
        # only token i goes through the network; rows 0..i-1 of the cache hold
        # the keys / values of the prefix
        xi = x[:, i].contiguous()
        self.backbone.run_decode(xi, n_fwd * B, L, i, mod, kv, step_logits,
                                 ids_period=B)
        la, lda = step_logits.data_ptr(), ld
        lb, ldb = step_logits.data_ptr() + B * ld * 2, ld
      else:
        self.backbone.run(x, n_fwd * B, L, mod, buf, ids_period=B * L)
        la, lda = buf.data_ptr() + i * ld * 2, L * ld
        lb, ldb = buf.data_ptr() + (B * L + i) * ld * 2, L * ld
      noise = None
      if self.noise_fn is not None:
        noise = self.noise_fn(i).to(self.device, torch.float32).contiguous()
      self._launch_sampler(
        logits_a=la, ld_a=lda, logits_b=lb if two else None,
        ld_b=ldb if two else 0, noise=noise, xt=y, xs=y, rows=B, seq_len=1,
        guidance=_lib.GUIDE_CFG if two else _lib.GUIDE_NONE,
        gamma=guidance.gamma if two else 1.0, diffusion=_lib.AR,
        select=_lib.SELECT_GUMBEL, step=i, q_out=q_next,
        compute_f64=self._use_float64 and method is None and q_next is None)
      if q_next is not None:
        thr = float(self.config.eval.get('low_confidence_threshold', 0.3))
        check(lib.ddg_low_confidence_select(
          ptr(q_next), V, B, V, thr, ptr(noise), V if noise is not None else 0,
          _lib.SELECT_GUMBEL, self.seed + 7919 * self._call_count, 0, i, 0, 0, -1, ptr(y),
          stream_ptr()))
      x[:, i + 1] = y
    return x


class _StepGraph:
  """One reverse step of the per-step API (`_ddpm_denoise` / `_cfg_denoise`)
  as a captured CUDA graph.  Inputs are copied into static buffers, the Philox
  key / step index are read from device memory, so one graph serves every call
  with the same shapes."""

  def __init__(self, model, kind, B, L, gamma):
    self.m, self.kind, self.B, self.L, self.gamma = model, kind, B, L, gamma
    self.two = kind == 'cfg2'
    self.n_fwd = 2 if self.two else 1
    dev, n = model.device, self.n_fwd * B
    self.xt = torch.empty(B, L, dtype=torch.int64, device=dev)
    self.xs = torch.empty(B, L, dtype=torch.int64, device=dev)
    self.sigma = torch.zeros(n, dtype=torch.float32, device=dev)
    self.labels = (torch.zeros(n, dtype=torch.int64, device=dev)
                   if kind != 'ddpm' else None)
    self.mct = torch.empty(B, dtype=torch.float32, device=dev)
    self.mcs = torch.empty(B, dtype=torch.float32, device=dev)
    self.step = torch.zeros(1, dtype=torch.int32, device=dev)
    self.seed = torch.zeros(1, dtype=torch.int64, device=dev)
    self.logits = model.backbone.alloc_logits(n * L)
    self.stats = model.backbone.alloc_stats(n * L)
    self.graph = None
    self.ws_generation = -1

  def logits_view(self, i):
    bb = self.m.backbone
    return self.logits.view(self.n_fwd, self.B, self.L, bb.vocab_ld)[
      i, :, :, :self.m.vocab_size]

  def _body(self):
    m, B, L = self.m, self.B, self.L
    mod = ModSpec(n_mod=self.n_fwd * B, seqs_per_mod=1, sigma=self.sigma,
                  sigma_row_stride=1, labels=self.labels)
    m.backbone.run(self.xt, self.n_fwd * B, L, mod, self.logits,
                   ids_period=B * L, stats_out=self.stats,
                   stats_mask_col=m._stats_mask_col)
    ld = m.backbone.vocab_ld
    m._launch_sampler(
      logits_a=self.logits.data_ptr(), ld_a=ld,
      logits_b=(self.logits.data_ptr() + B * L * ld * 2) if self.two else None,
      ld_b=ld if self.two else 0, xt=self.xt, xs=self.xs, mc_t=self.mct,
      mc_s=self.mcs, rows=B * L, seq_len=L, mc_seq_stride=1,
      step_ptr=self.step, seed_ptr=self.seed,
      stats=(self.stats, 0, B * L if self.two else 0),
      guidance=_lib.GUIDE_CFG if self.two else _lib.GUIDE_NONE,
      gamma=self.gamma if self.two else 1.0)

  def run(self, xt, sigma, labels, mc_t, mc_s):
    m, B = self.m, self.B
    self.xt.copy_(xt, non_blocking=True)
    sigma = sigma.to(m.device, torch.float32).reshape(-1)
    self.sigma.copy_(sigma.repeat(self.n_fwd) if sigma.numel() == B
                     else sigma.expand(self.n_fwd * B))
    if self.labels is not None:
      self.labels.copy_(labels.to(m.device, torch.int64).reshape(-1))
    self.mct.copy_(m._mc(mc_t).expand(B))
    self.mcs.copy_(m._mc(mc_s).expand(B))
    self.step.fill_(int(m._step_index))
    self.seed.fill_(m.seed + 7919 * m._call_count)
    m._step_index += 1          # successive API calls draw fresh noise
    if self.graph is not None and self.ws_generation != m.backbone.ws_generation:
      self.graph = None
    if self.graph is None:
      self._body()              # eager (also warms lazy attributes), then capture
      torch.cuda.current_stream().synchronize()
      out = self.xs.clone()
      self.graph = _capture(self._body)
      self.ws_generation = m.backbone.ws_generation
      return out
    self.graph.replay()
    return self.xs.clone()


class _GraphLoop:
  """One reverse step captured in a CUDA graph.  Per-step scalars live in
  device tables indexed by a device-side step counter that the last node of
  the graph increments, so the same graph is replayed for every step."""

  def __init__(self, model: Diffusion, B, L, method, gamma, key):
    self.key = key
    self.model = model
    self.B, self.L = B, L
    self.method = method
    self.gamma = gamma
    self.two = method == 'cfg' and gamma not in (0.0, 1.0)
    self.n_fwd = 2 if self.two else 1
    dev = model.device
    self.xt = torch.empty(B, L, dtype=torch.int64, device=dev)
    self.step = torch.zeros(1, dtype=torch.int32, device=dev)
    self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)  # Philox key
    self.logits = model.backbone.alloc_logits(self.n_fwd * B * L)
    self.stats = model.backbone.alloc_stats(self.n_fwd * B * L)
    self.labels = (torch.zeros(self.n_fwd, dtype=torch.int64, device=dev)
                   if method == 'cfg' else None)
    # D-CBG (diffusion.py:1511-1624): denoiser forward, classifier term (first order: one
    # forward + backward; exact: every single-token edit), guided sampler -- all in the
    # captured step, the term in a persistent buffer
    self.classifier = self.term = None
    if method == 'cbg':
      g = model.config.guidance
      self.classifier = model.classifier_model
      self.condition = int(g.condition)
      self.use_approx = bool(g.get('use_approx', False))
      ldv = (model.vocab_size + 7) // 8 * 8 if self.use_approx else model.vocab_size
      self.term = torch.empty(B * L, ldv, dtype=torch.float32, device=dev)
    self.tables = None
    self.graph = None
    self.ws_generation = -1
    self.steps = 0
    # Logits caches (both result-identical to what the reference computes):
    # 'seq'   Without time conditioning the logits of a sequence are a function of its tokens
    #         alone, and absorbing diffusion changes L / steps tokens per sequence per step on
    #         average (0.128 at L = 128, 1000 steps): most sequences are untouched by a step.
    #         The sampler flags the sequences it changes, `ddg_seq_compact` turns the flags
    #         into the next step's dynamic batch, and the forward runs on those sequences
    #         only, writing their rows of the persistent logits buffer.  The reference does
    #         this for the whole batch at once (sampling.use_cache: reuse when NO token
    #         changed, diffusion.py:1371-1374); same results, the forward being deterministic
    #         per sequence.
    # 'batch' With time conditioning the reference's `sampling.use_cache` is NOT a no-op: while
    #         no token of the batch changes it keeps the log-probs computed at an OLDER sigma
    #         (diffusion.py:1371-1374 with :1386-1390, 1448-1470).  Reproduced exactly: the
    #         forward of a step runs on all sequences if any token changed in the previous
    #         step and on none otherwise (device-side count, no host sync).
    self.cached = None
    if (model.logits_cache and method is None and not model.time_conditioning
        and model.diffusion == 'absorbing_state' and L % 32 == 0):
      self.cached = 'seq'
    elif model.config.sampling.get('use_cache', False) and model.time_conditioning:
      self.cached = 'batch'
    if self.cached:
      self.changed = torch.ones(B, dtype=torch.int32, device=dev)
      self.seq_map = (torch.zeros(B, dtype=torch.int32, device=dev)
                      if self.cached == 'seq' else None)
      self.n_live = torch.zeros(1, dtype=torch.int32, device=dev)

  def _body(self):
    m, B, L = self.model, self.B, self.L
    sig, mct, mcs = self.tables
    mod = ModSpec(n_mod=self.n_fwd, seqs_per_mod=B, sigma=sig,
                  sigma_row_stride=0, sigma_step_stride=1, step_ptr=self.step,
                  labels=self.labels)
    live = None
    if self.cached:    # flags of the previous step (all set at step 0) -> this step's batch
      check(lib.ddg_seq_compact(ptr(self.changed), B, ptr(self.seq_map), ptr(self.n_live),
                                self.n_fwd * B if self.cached == 'batch' else 0, stream_ptr()))
      live = (self.seq_map, self.n_live)
    m.backbone.run(self.xt, self.n_fwd * B, L, mod, self.logits,
                   ids_period=B * L, stats_out=self.stats,
                   stats_mask_col=m._stats_mask_col, live=live)
    ld = m.backbone.vocab_ld
    term = None
    if self.method == 'cbg':
      # this step's sigma from the device table (one shared conditioning row)
      sigma_now = sig.index_select(0, self.step).expand(B)
      term = m._classifier_term(self.classifier, self.xt, sigma_now, self.condition,
                                self.use_approx, raw_ok=True, out=self.term,
                                live=self.n_live if self.cached else None)
    guide = (_lib.GUIDE_CFG if self.two else
             _lib.GUIDE_CBG if term is not None else _lib.GUIDE_NONE)
    m._launch_sampler(
      logits_a=self.logits.data_ptr(), ld_a=ld,
      logits_b=(self.logits.data_ptr() + B * L * ld * 2) if self.two else None,
      ld_b=ld if self.two else 0, clf=term, xt=self.xt, xs=self.xt, mc_t=mct, mc_s=mcs,
      rows=B * L, seq_len=L, mc_seq_stride=0, mc_step_stride=1,
      step_ptr=self.step, seed_ptr=self.seed_dev,
      stats=(self.stats, 0, B * L if self.two else 0),
      guidance=guide, gamma=self.gamma if guide != _lib.GUIDE_NONE else 1.0,
      seq_changed=self.changed if self.cached else None)
    check(lib.ddg_step_advance(ptr(self.step), stream_ptr()))

  def prepare(self, xt, sigma_t, mc_t, mc_s, cond):
    m = self.model
    steps = sigma_t.numel()
    dev = m.device
    sig = sigma_t if m.time_conditioning else torch.zeros_like(sigma_t)
    if self.tables is None or self.steps != steps:
      self.tables = tuple(torch.empty(steps, dtype=torch.float32, device=dev)
                          for _ in range(3))
      self.steps = steps
      self.graph = None
    for dst, src in zip(self.tables, (sig, mc_t, mc_s)):
      dst.copy_(src.to(torch.float32))
    if self.method == 'cfg':
      num_classes = m.config.data.num_classes
      c = int(cond[0].item()) if cond is not None else 0
      if self.two:
        lab = [c, num_classes]
      else:
        lab = [num_classes if self.gamma == 0.0 else c]
      self.labels.copy_(torch.tensor(lab, dtype=torch.int64))
    self.xt.copy_(xt)
    self.step.zero_()
    self.seed_dev.fill_(m.seed + 7919 * m._call_count)
    if self.graph is not None and self.ws_generation != m.backbone.ws_generation:
      self.graph = None    # workspace buffers were reallocated since capture
    if self.cached:
      self.changed.fill_(1)      # new prior: every sequence needs its first forward
    if self.graph is None:
      # one eager step first (lazy attribute / module setup), then capture
      self._body()
      torch.cuda.synchronize()
      self.reset(xt)
      self.graph = _capture(self._body)
      self.ws_generation = m.backbone.ws_generation
      self.reset(xt)

  def reset(self, xt):
    """Back to step 0 of the schedule tables with a fresh prior."""
    self.xt.copy_(xt)
    self.step.zero_()
    if self.cached:
      self.changed.fill_(1)

  def replay(self, n):
    for _ in range(n):
      self.graph.replay()

  def run(self, xt, sigma_t, mc_t, mc_s, cond):
    self.prepare(xt, sigma_t, mc_t, mc_s, cond)
    self.replay(self.steps)
    return self.xt.clone()
