"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np
import torch

import ddg_b200
from golden import cases
from oracle import ref_sampler

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOPK = 8


def load_golden(name):
  return dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')))


def case_config(case):
  return ddg_b200.config.to_config(cases.config_dict(case))


def case_state(case):
  cfg = case_config(case)
  tok = cases.Tok(case.vocab)
  sd = ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=case.seed)
  csd = None
  if case.guidance == 'cbg':
    csd = ddg_b200.weights.random_state_dict(cfg, tok.vocab_size,
                                             seed=case.seed + 1,
                                             classifier=True)
  return cfg, tok, sd, csd


def make_oracle(case, gpu_rounding=False):
  cfg, tok, sd, csd = case_state(case)
  return ref_sampler.OracleSampler(
    sd, vocab_size=tok.vocab_size, mask_index=tok.mask_token_id,
    parameterization=case.parameterization, diffusion=case.diffusion,
    n_heads=case.n_heads, time_conditioning=case.time_conditioning,
    subs_masking=case.subs_masking, T=case.T, num_classes=case.num_classes,
    gpu_rounding=gpu_rounding, argmax_sampling=case.argmax_sampling,
    use_float64=case.use_float64,
    low_confidence=case.low_confidence_threshold if case.low_confidence else None,
    clf_sd=csd, clf_n_heads=case.n_heads)


def check_summary(name, t, gold, big, rtol, atol):
  """Compares tensor t with the fixture entry `name` (full tensor, or the
  top-k / sum / strided-column summary stored for the 30522-vocab cases)."""
  t = t.detach().float().cpu()
  if not big:
    np.testing.assert_allclose(t.numpy(), gold[name], rtol=rtol, atol=atol,
                               err_msg=name)
    return
  top = t.topk(TOPK, dim=-1)
  np.testing.assert_allclose(top.values.numpy(), gold[name + '_topv'],
                             rtol=rtol, atol=atol, err_msg=name + '_topv')
  np.testing.assert_allclose(t.double().sum(-1).numpy(), gold[name + '_sum'],
                             rtol=max(rtol, 1e-5), atol=atol,
                             err_msg=name + '_sum')
  np.testing.assert_allclose(t[..., ::997].numpy(), gold[name + '_cols'],
                             rtol=rtol, atol=atol, err_msg=name + '_cols')


def near_tie_mask(score_fn, ids_a, ids_b, tol=1e-5):
  """True where ids differ only because the two candidates' scores tie to
  within `tol` (relative)."""
  sa, sb = score_fn(ids_a), score_fn(ids_b)
  return (sa - sb).abs() <= tol * sa.abs().clamp_min(1e-30)


# ----------------------------------------------------------------- GPU side
def philox_uniform(seed, step, rows, vocab):
  """CPU regeneration of the sampler's in-kernel Philox4x32-10 stream
  (csrc/ddg_common.cuh): key=(seed_lo, seed_hi), counter=(j/4, row_lo, row_hi,
  step), U = (word >> 8) * 2^-24.  Returns (rows, vocab) float32."""
  M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
  W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
  nj4 = (vocab + 3) // 4
  c0 = np.broadcast_to(np.arange(nj4, dtype=np.uint32)[None, :], (rows, nj4)).copy()
  r = np.arange(rows, dtype=np.uint64)[:, None]
  c1 = np.broadcast_to((r & np.uint64(0xFFFFFFFF)).astype(np.uint32), (rows, nj4)).copy()
  c2 = np.broadcast_to((r >> np.uint64(32)).astype(np.uint32), (rows, nj4)).copy()
  c3 = np.full((rows, nj4), np.uint32(step & 0xFFFFFFFF), dtype=np.uint32)
  k0 = np.uint32(seed & 0xFFFFFFFF)
  k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = M0 * c0.astype(np.uint64)
      p1 = M1 * c2.astype(np.uint64)
      hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
      hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
      c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
      k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
      k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
  words = np.stack([c0, c1, c2, c3], axis=-1).reshape(rows, nj4 * 4)[:, :vocab]
  return ((words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))


def sampler_kernel(*, logits_a, xt, mc_t, mc_s, vocab, seq_len, diffusion,
                   logits_b=None, clf=None, noise=None, guidance='none',
                   gamma=1.0, subs=False, mask_logit=False, mask_index=0,
                   round_bf16=False, select='race', fast_math=False, seed=0,
                   step=0, want_q=False, stats_a=0, stats_b=0, stats_ld=0,
                   stats_chunks=0, race_group=0, compute_f64=False,
                   clf_round_bf16=False):
  """Direct call of ddg_sample_step through the C ABI on torch CUDA tensors.
  logits (rows, ld) f32 or bf16; xt (rows,) int64; mc_* scalars or (n_seq,)."""
  from ddg_b200 import _lib
  dev = logits_a.device
  rows = xt.numel()
  xt = xt.reshape(-1).to(dev, torch.int64).contiguous()
  xs = torch.empty_like(xt)
  q = torch.empty(rows, vocab, dtype=torch.float32, device=dev) if want_q else None
  mct = torch.as_tensor(mc_t, dtype=torch.float32).reshape(-1).to(dev).contiguous()
  mcs = torch.as_tensor(mc_s, dtype=torch.float32).reshape(-1).to(dev).contiguous()
  dcode = {'absorbing_state': _lib.ABSORBING, 'uniform': _lib.UNIFORM, 'ar': _lib.AR}[diffusion]
  gcode = {'none': _lib.GUIDE_NONE, None: _lib.GUIDE_NONE, 'cfg': _lib.GUIDE_CFG, 'cbg': _lib.GUIDE_CBG}[guidance]
  scode = {'race': _lib.SELECT_RACE, 'argmax': _lib.SELECT_ARGMAX, 'gumbel': _lib.SELECT_GUMBEL}[select]
  if noise is not None:
    noise = noise.reshape(rows, vocab).to(dev, torch.float32).contiguous()
  ld_clf = vocab
  if clf is not None:
    if clf.dim() == 2 and clf.shape[0] == rows and clf.shape[1] > vocab:   # rows padded by the caller
      clf = clf.to(dev, torch.float32).contiguous()
      ld_clf = clf.stride(0)
    else:
      clf = clf.reshape(rows, vocab).to(dev, torch.float32).contiguous()
  a = _lib.SampleStepArgs(
    logits_a=logits_a.data_ptr(), logits_b=_lib.ptr(logits_b), clf=_lib.ptr(clf),
    noise=_lib.ptr(noise), xt=xt.data_ptr(), mc_t=mct.data_ptr(),
    mc_s=mcs.data_ptr(), step_ptr=0, seed_ptr=0, xs=xs.data_ptr(),
    q_out=_lib.ptr(q), rows=rows, ld_a=logits_a.stride(0),
    ld_b=logits_b.stride(0) if logits_b is not None else 0, ld_clf=ld_clf,
    ld_noise=vocab, ld_q=vocab, seq_len=seq_len, vocab=vocab,
    mc_seq_stride=1 if mct.numel() > 1 else 0, mc_step_stride=0,
    logits_dtype=_lib.DTYPE_CODE[logits_a.dtype], diffusion=dcode,
    guidance=gcode, subs=int(subs), mask_logit=int(mask_logit),
    mask_index=mask_index, round_bf16=int(round_bf16), select=scode,
    fast_math=int(fast_math), gamma=float(gamma), seed=int(seed),
    step=int(step), stats_chunks=stats_chunks, stats_a=stats_a, stats_b=stats_b,
    stats_ld=stats_ld, race_group=race_group, compute_f64=int(compute_f64),
    clf_round_bf16=int(clf_round_bf16))
  _lib.check(_lib.lib.ddg_sample_step(a, _lib.stream_ptr()))
  torch.cuda.synchronize()
  return xs, q


def low_confidence_kernel(q, threshold, *, noise=None, select='race', xt=None,
                          carry_mask_index=-1, seed=0, step=0):
  """Direct call of ddg_low_confidence_select on a (rows, V) fp32 posterior."""
  from ddg_b200 import _lib
  dev = q.device
  rows, V = q.shape
  q = q.contiguous()
  xs = torch.empty(rows, dtype=torch.int64, device=dev)
  if noise is not None:
    noise = noise.reshape(rows, V).to(dev, torch.float32).contiguous()
  if xt is not None:
    xt = xt.reshape(-1).to(dev, torch.int64).contiguous()
  scode = {'race': _lib.SELECT_RACE, 'argmax': _lib.SELECT_ARGMAX}[select]
  _lib.check(_lib.lib.ddg_low_confidence_select(
    q.data_ptr(), V, rows, V, float(threshold), _lib.ptr(noise), V if noise is not None else 0,
    scode, int(seed), 0, int(step), 0, _lib.ptr(xt), int(carry_mask_index), xs.data_ptr(),
    _lib.stream_ptr()))
  torch.cuda.synchronize()
  return xs


def low_confidence_mask(q, threshold):
  """The reference's kept set (diffusion.py:1739-1746) for a (..., V) posterior."""
  sp, si = q.sort(dim=-1, stable=True)
  ms = sp.cumsum(dim=-1) <= threshold
  mask = torch.zeros_like(ms)
  mask.scatter_(dim=-1, index=si, src=ms)
  return mask


def race_scores(q, U):
  """Reference selection statistic p / E (diffusion.py:1752-1755)."""
  return q / (1e-10 - (U + 1e-10).log())


def assert_ids_match(xs_gpu, xs_ref, q_ref, U, what, tol=2e-5):
  """Bit-exact ids except documented near-ties: where they differ, the
  reference's own selection scores of the two candidates must agree to `tol`
  relative (fp32 exp/log implementations differ in the last ulp)."""
  xs_gpu = xs_gpu.cpu().reshape(xs_ref.shape)
  diff = xs_gpu != xs_ref
  n = int(diff.sum())
  if n == 0:
    return 0
  s = race_scores(q_ref.double(), U.double()) if U is not None else q_ref.double()
  s_ref = s.gather(-1, xs_ref[..., None])[..., 0][diff]
  s_gpu = s.gather(-1, xs_gpu[..., None])[..., 0][diff]
  rel = (s_ref - s_gpu).abs() / s_ref.abs().clamp_min(1e-300)
  assert bool((rel <= tol).all()), (
    f'{what}: {n} ids differ, worst relative score gap {rel.max().item():.3e}')
  assert n <= max(2, xs_ref.numel() // 50), f'{what}: too many near-ties ({n})'
  return n


def assert_close_frac(a, b, rtol, atol, loose_frac=0.0, loose_rtol=0.0):
  """|a-b| <= atol + rtol|b| everywhere, except that a fraction `loose_frac`
  of the entries may deviate up to loose_rtol|b| (+atol)."""
  a, b = a.detach().float().cpu(), b.detach().float().cpu()
  err = (a - b).abs()
  bad = err > atol + rtol * b.abs()
  if not bool(bad.any()):
    return
  frac = bad.float().mean().item()
  assert frac <= loose_frac, f'{frac:.2e} of entries outside rtol={rtol}'
  worst = (err[bad] / (atol + b.abs()[bad])).max().item()
  assert bool((err[bad] <= atol + loose_rtol * b.abs()[bad]).all()), (
    f'loose entries exceed rtol={loose_rtol}: worst {worst:.3e}')


def make_stats(logits, vocab, mask_col=-1):
  """Torch restatement of the vocab-GEMM epilogue's log-softmax partials
  (ddg_gemm_args.stats): per 128-column chunk (max, sum 2^((v-max) log2 e)),
  laid out (chunks, rows, 2) fp32."""
  rows = logits.shape[0]
  v = logits[:, :vocab].float()
  if mask_col >= 0:
    v = v.clone()
    v[:, mask_col] = float('-inf')
  chunks = (vocab + 127) // 128
  pad = chunks * 128 - vocab
  if pad:
    v = torch.cat([v, torch.full((rows, pad), float('-inf'), device=v.device)], 1)
  v = v.view(rows, chunks, 128)
  m = v.amax(-1)
  s = torch.exp2((v - m[..., None]) * 1.4426950408889634).sum(-1)
  return torch.stack([m, s], -1).permute(1, 0, 2).contiguous()


# ----------------------------------------------------------------- reference traces
def gpu_diffusion(case):
  cfg, tok, sd, csd = case_state(case)
  model = ddg_b200.Diffusion(cfg, tok, device='cuda').load_state_dict(sd)
  if csd is not None:
    model.set_classifier(ddg_b200.Classifier(cfg, tok, device='cuda').load_state_dict(csd))
  return model, cfg, tok


def _full(gold, key, big):
  """Tensor stored under `key` (full) or None when only a summary was kept."""
  return None if big else torch.from_numpy(gold[key]).float()


def check_against_reference_trace(case, gold, *, min_exact, max_rel_lp, q_med, q_max,
                                  id_frac):
  """Replays the reference's trace `gold` (per-step xt, scalars, log-probs,
  posterior, ids) through the CUDA per-step API with the same injected noise.

  * bf16 log-probs: two pipelines that round at the same points agree bit for bit
    wherever accumulation-order noise did not flip a rounding -> a share
    `min_exact` of the entries must be EXACTLY equal, the rest within `max_rel_lp`
    (a few bf16 ulps);
  * posterior q_xs: median / max relative deviation;
  * ids: may differ only on a small fraction (near-ties of the race under
    bf16-rounded probabilities).
  Returns the measured statistics."""
  model, cfg, tok = gpu_diffusion(case)
  model.return_q_xs = True
  big = case.vocab == 'bert'
  B, L, V = case.batch, case.length, tok.vocab_size
  stats = dict(exact=[], rel_lp=[], q_med=[], q_max=[], n_ids=0, n_diff=0)
  for i in range(case.steps):
    U = cases.step_noise(case, i)
    model.noise_fn = lambda _i, U=U: U
    xt = torch.from_numpy(gold[f's{i}_xt']).long().to('cuda')
    sc = gold[f's{i}_scalars']
    sigma_t = torch.full((B,), float(sc[0]), device='cuda')
    mc_t = torch.full((B, 1, 1), float(sc[1]), device='cuda')
    mc_s = torch.full((B, 1, 1), float(sc[2]), device='cuda')
    if case.guidance is None:
      xs, q, cache = model._ddpm_denoise(xt, sigma_t, mc_t, mc_s)
    elif case.guidance == 'cfg':
      cond = torch.full((B,), case.condition, dtype=torch.long, device='cuda')
      xs, q, cache = model._cfg_denoise(cond, case.gamma, xt, sigma_t, mc_t, mc_s)
    else:
      xs, q, cache = model._cbg_denoise(
        case.condition, case.gamma, model.classifier_model, xt, sigma_t, mc_t, mc_s,
        use_approx=case.use_approx)
    torch.cuda.synchronize()
    for k, v in cache.items():
      if v is None or not k.startswith('log_x_theta'):
        continue
      # the API caches logits; the reference caches their (bf16) log-softmax
      lp = ref_sampler.log_probs_from_logits(
        v.float().cpu(), xt.cpu(), parameterization=case.parameterization,
        mask_index=tok.mask_token_id, subs_masking=case.subs_masking,
        lp_dtype=torch.bfloat16)
      if big:
        g = torch.from_numpy(gold[f's{i}_{k}_cols']).float()
        lp = lp[..., ::997]
      else:
        g = torch.from_numpy(gold[f's{i}_{k}']).float()
      fin = torch.isfinite(g) & (g > -1e5) & (g != 0)
      rel = ((lp - g).abs() / g.abs())[fin]
      stats['exact'].append((rel == 0).float().mean().item())
      stats['rel_lp'].append(rel.max().item())
    qc = q.float().cpu().view(B, L, V)
    if big:
      g = torch.from_numpy(gold[f's{i}_q_xs_cols']).float()
      qc = qc[..., ::997]
    else:
      g = torch.from_numpy(gold[f's{i}_q_xs']).float()
    pos = g > 1e-9
    rel = ((qc - g).abs() / g.clamp_min(1e-30))[pos]
    stats['q_med'].append(rel.median().item())
    stats['q_max'].append(rel.max().item())
    gx = torch.from_numpy(gold[f's{i}_xs']).long()
    stats['n_ids'] += gx.numel()
    stats['n_diff'] += int((xs.cpu() != gx).sum())
  summary = dict(exact=min(stats['exact']), rel_lp=max(stats['rel_lp']),
                 q_med=max(stats['q_med']), q_max=max(stats['q_max']),
                 n_diff=stats['n_diff'], n_ids=stats['n_ids'])
  print(f'[reference-on-gpu] {case.name}: {summary}')
  assert summary['exact'] >= min_exact, summary
  assert summary['rel_lp'] <= max_rel_lp, summary
  assert summary['q_med'] <= q_med and summary['q_max'] <= q_max, summary
  assert summary['n_diff'] <= max(2, int(id_frac * summary['n_ids'])), summary
  return summary
