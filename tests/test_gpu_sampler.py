"""GPU parity of the fused sampler kernel (ddg_sample_step, C ABI).

Tolerances (stated once, used below):
  * ids: bit-exact against the reference / oracle on identical logits and
    injected noise, except near-ties (selection scores within 2e-5 relative);
  * q_xs: |dq| <= 1e-6 + 2e-4 |q|  (fp32 exp/log implementation differences).
"""
import numpy as np
import pytest
import torch

import helpers
from golden import cases
from oracle import ref_dit, ref_sampler

pytestmark = pytest.mark.gpu

DIFFUSION_CASES = [c.name for c in cases.CASES if c.parameterization != 'ar' and not c.use_cache]


def _oracle_logits(case, oracle, xt, sigma_t, cond):
  """Raw fp32 backbone logits from the CPU oracle for one step."""
  sigma = sigma_t if case.time_conditioning else torch.zeros_like(sigma_t)
  return ref_dit.dit_forward(oracle.sd, xt, sigma, cond, n_heads=case.n_heads)


@pytest.mark.parametrize('name', DIFFUSION_CASES)
def test_kernel_replays_reference_goldens(name):
  """Feeds the kernel the same fp32 logits and uniforms the reference saw and
  requires the reference's own ids and posterior (tests/golden/*.npz)."""
  case = cases.CASE_BY_NAME[name]
  gold = helpers.load_golden(name)
  big = case.vocab == 'bert'
  oracle = helpers.make_oracle(case)
  tok = cases.Tok(case.vocab)
  V, B, L = tok.vocab_size, case.batch, case.length
  dev = torch.device('cuda')
  near = 0
  for i in range(case.steps):
    xt = torch.from_numpy(gold[f's{i}_xt']).long()
    sc = gold[f's{i}_scalars']
    sigma_t = torch.full((B,), float(sc[0]))
    U = cases.step_noise(case, i)
    kw = dict(xt=xt, mc_t=float(sc[1]), mc_s=float(sc[2]), vocab=V, seq_len=L,
              diffusion=case.diffusion, noise=U,
              subs=case.parameterization == 'subs',
              mask_logit=case.parameterization == 'subs' or case.subs_masking,
              mask_index=tok.mask_token_id, round_bf16=False, fast_math=False,
              select='argmax' if case.argmax_sampling else 'race', want_q=True,
              compute_f64=case.use_float64)
    if case.guidance == 'cfg':
      cond = torch.full((B,), case.condition, dtype=torch.long)
      la = lb = None
      if case.gamma != 0.0:
        la = _oracle_logits(case, oracle, xt, sigma_t, cond)
      if case.gamma != 1.0:
        lb = _oracle_logits(case, oracle, xt, sigma_t,
                            torch.full_like(cond, case.num_classes))
      if la is None or lb is None:
        one = la if la is not None else lb
        xs, q = helpers.sampler_kernel(
          logits_a=one.reshape(B * L, V).to(dev).contiguous(), **kw)
      else:
        xs, q = helpers.sampler_kernel(
          logits_a=la.reshape(B * L, V).to(dev).contiguous(),
          logits_b=lb.reshape(B * L, V).to(dev).contiguous(),
          guidance='cfg', gamma=case.gamma, **kw)
    elif case.guidance == 'cbg':
      la = _oracle_logits(case, oracle, xt, sigma_t, None)
      term_fn = (ref_sampler.cbg_classifier_term_approx if case.use_approx
                 else ref_sampler.cbg_classifier_term_exact)
      term = term_fn(oracle.clf_log_probs, xt, sigma_t, V, case.condition)
      xs, q = helpers.sampler_kernel(
        logits_a=la.reshape(B * L, V).to(dev).contiguous(), clf=term,
        guidance='cbg', gamma=case.gamma, **kw)
    else:
      la = _oracle_logits(case, oracle, xt, sigma_t, None)
      xs, q = helpers.sampler_kernel(
        logits_a=la.reshape(B * L, V).to(dev).contiguous(), **kw)
    q = q.view(B, L, V)
    helpers.check_summary(f's{i}_q_xs', q, gold, big, rtol=2e-4, atol=1e-6)
    gold_xs = torch.from_numpy(gold[f's{i}_xs']).long()
    # the 30522-vocab fixtures store q_xs as a summary only: score near-ties
    # with the (just verified) kernel posterior there
    q_ref = q.cpu() if big else torch.from_numpy(gold[f's{i}_q_xs'])
    if case.low_confidence:
      # eval.low_confidence_sampling: the draw is redone on the kept set
      absorbing = case.diffusion == 'absorbing_state'
      xs = helpers.low_confidence_kernel(
        q.reshape(B * L, V), case.low_confidence_threshold, noise=U,
        select='argmax' if case.argmax_sampling else 'race', xt=xt if absorbing else None,
        carry_mask_index=tok.mask_token_id if absorbing else -1)
      q_ref = q_ref.masked_fill(~helpers.low_confidence_mask(q_ref, case.low_confidence_threshold), 0.0)
      if absorbing:   # carried-over rows: compare through their one-hot posterior
        keep = xt != tok.mask_token_id
        q_ref[keep] = torch.nn.functional.one_hot(xt[keep], V).float()
    near += helpers.assert_ids_match(
      xs.view(B, L), gold_xs, q_ref, None if case.argmax_sampling else U,
      f'{name} step {i}')
  print(f'{name}: near-ties {near}')


def _rand_logits(rows, V, ld, dtype, seed, scale=3.0):
  g = torch.Generator().manual_seed(seed)
  full = torch.zeros(rows, ld, dtype=torch.float32)
  full[:, :V] = torch.randn(rows, V, generator=g) * scale
  return full.to(dtype)


MODES = [
  ('uniform', 'none'), ('uniform', 'cfg'), ('uniform', 'cbg'),
  ('absorbing_state', 'none'), ('absorbing_state', 'cfg'),
  ('absorbing_state', 'cbg'),
]


@pytest.mark.parametrize('diffusion,guidance', MODES)
@pytest.mark.parametrize('V,ld', [(30522, 30528), (30522, 30522), (35, 64), (1000, 1000)])
@pytest.mark.parametrize('round_bf16', [True, False])
def test_kernel_vs_oracle_bf16_logits(diffusion, guidance, V, ld, round_bf16):
  """bf16 logits (what the DiT head emits), every mode, aligned / unaligned
  leading dims, with and without the reference's GPU roundings."""
  B, L = 2, 6
  rows = B * L
  dev = torch.device('cuda')
  mask = 4 if V < 100 else 103
  subs = diffusion == 'absorbing_state'
  la = _rand_logits(rows, V, ld, torch.bfloat16, 1)
  lb = _rand_logits(rows, V, ld, torch.bfloat16, 2)
  g = torch.Generator().manual_seed(3)
  U = torch.rand(B, L, V, generator=g)
  clf = torch.randn(B, L, V, generator=g) * 0.5 - 1.0
  if subs:
    xt = torch.full((B, L), mask, dtype=torch.long)
    xt[0, :2] = torch.tensor([7, 9])      # two carried-over tokens
  else:
    xt = torch.randint(0, V, (B, L), generator=g)
  mc_t = torch.tensor([0.61, 0.37])
  mc_s = torch.tensor([0.59, 0.33])
  lp_dtype = torch.bfloat16 if round_bf16 else None
  param = 'subs' if subs else 'd3pm'

  def lps(l):
    return ref_sampler.log_probs_from_logits(
      l[:, :V].float().view(B, L, V), xt, parameterization=param,
      mask_index=mask, lp_dtype=lp_dtype)
  kw = dict(diffusion=diffusion, mask_index=mask, lp_dtype=lp_dtype)
  mct3, mcs3 = mc_t[:, None, None], mc_s[:, None, None]
  if guidance == 'none':
    xs_ref, q_ref = ref_sampler.ddpm_update(lps(la), xt, mct3, mcs3, U, **kw)
  elif guidance == 'cfg':
    xs_ref, q_ref = ref_sampler.cfg_update(lps(la), lps(lb), 2.0, xt, mct3, mcs3, U, **kw)
  else:
    # round_bf16 runs take the exact classifier term: a bf16 tensor in the reference,
    # so gamma * term is itself rounded to bf16 (diffusion.py:1607, 1612)
    xs_ref, q_ref = ref_sampler.cbg_update(lps(la), clf, 1.5, xt, mct3, mcs3, U,
                                           clf_dtype=lp_dtype, **kw)
  xs, q = helpers.sampler_kernel(
    logits_a=la.to(dev), logits_b=lb.to(dev) if guidance == 'cfg' else None,
    clf=clf if guidance == 'cbg' else None, noise=U, xt=xt, mc_t=mc_t,
    mc_s=mc_s, vocab=V, seq_len=L, diffusion=diffusion, guidance=guidance,
    gamma=2.0 if guidance == 'cfg' else 1.5, subs=subs, mask_logit=subs,
    mask_index=mask, round_bf16=round_bf16, fast_math=False, want_q=True,
    clf_round_bf16=round_bf16 and guidance == 'cbg')
  q = q.view(B, L, V).cpu()
  # with the bf16 roundings on, an fp32 last-ulp difference can flip a bf16
  # rounding (0.4 % of that element): allow <= 0.1 % of entries to move 1 bf16 ulp
  helpers.assert_close_frac(q, q_ref, rtol=3e-4, atol=1e-6,
                            loose_frac=1e-3 if round_bf16 else 0.0,
                            loose_rtol=2e-2)
  helpers.assert_ids_match(xs.view(B, L), xs_ref, q_ref, U,
                           f'{diffusion}/{guidance}/V{V}',
                           tol=2e-2 if round_bf16 else 2e-5)
  if subs:   # carry-over (diffusion.py:1414-1418)
    assert xs.view(B, L)[0, 0].item() == 7 and xs.view(B, L)[0, 1].item() == 9


@pytest.mark.parametrize('diffusion,guidance', MODES)
def test_philox_stream_matches_cpu_regeneration(diffusion, guidance):
  """In-kernel Philox mode == injected-noise mode fed the same stream
  regenerated on the CPU (pins the production RNG path bit for bit)."""
  B, L, V, ld = 2, 5, 30522, 30528
  rows = B * L
  dev = torch.device('cuda')
  subs = diffusion == 'absorbing_state'
  la = _rand_logits(rows, V, ld, torch.bfloat16, 5).to(dev)
  lb = _rand_logits(rows, V, ld, torch.bfloat16, 6).to(dev)
  g = torch.Generator().manual_seed(7)
  clf = torch.randn(B, L, V, generator=g) * 0.5
  xt = (torch.full((B, L), 103, dtype=torch.long) if subs
        else torch.randint(0, V, (B, L), generator=g))
  seed, step = 0x1234567890ABCDEF, 17
  U = torch.from_numpy(helpers.philox_uniform(seed, step, rows, V))
  common = dict(logits_a=la, logits_b=lb if guidance == 'cfg' else None,
                clf=clf if guidance == 'cbg' else None, xt=xt, mc_t=0.5,
                mc_s=0.05, vocab=V, seq_len=L, diffusion=diffusion,
                guidance=guidance, gamma=2.0, subs=subs, mask_logit=subs,
                mask_index=103, round_bf16=True, fast_math=False)
  xs_inj, _ = helpers.sampler_kernel(noise=U, **common)
  xs_phx, _ = helpers.sampler_kernel(noise=None, seed=seed, step=step, **common)
  assert torch.equal(xs_inj, xs_phx)
  xs_other, _ = helpers.sampler_kernel(noise=None, seed=seed, step=step + 1, **common)
  assert not torch.equal(xs_other, xs_phx)


def test_fast_math_agrees_with_careful_math():
  """Production (ex2/lg2 intrinsics) vs parity math: same ids except a few
  near-ties, on a full-size LM1B UDLM D-CFG tile."""
  B, L, V, ld = 8, 128, 30522, 30528
  rows = B * L
  dev = torch.device('cuda')
  g = torch.Generator(device=dev).manual_seed(1)
  la = (torch.randn(rows, ld, device=dev, generator=g) * 3).to(torch.bfloat16)
  lb = (torch.randn(rows, ld, device=dev, generator=g) * 3).to(torch.bfloat16)
  xt = torch.randint(0, V, (rows,), device=dev, generator=g)
  common = dict(logits_a=la, logits_b=lb, xt=xt, mc_t=0.5, mc_s=0.499,
                vocab=V, seq_len=L, diffusion='uniform', guidance='cfg',
                gamma=2.0, round_bf16=True, seed=9, step=3)
  a, _ = helpers.sampler_kernel(fast_math=False, **common)
  b, _ = helpers.sampler_kernel(fast_math=True, **common)
  frac = (a != b).float().mean().item()
  assert frac < 5e-3, frac


def test_full_size_properties():
  """BASELINE.json config 3 shape per GPU (64 x 128 x 30522): properties that
  do not need the CPU oracle."""
  B, L, V, ld = 64, 128, 30522, 30528
  rows = B * L
  dev = torch.device('cuda')
  g = torch.Generator(device=dev).manual_seed(2)
  la = (torch.randn(rows, ld, device=dev, generator=g) * 3).to(torch.bfloat16)
  lb = (torch.randn(rows, ld, device=dev, generator=g) * 3).to(torch.bfloat16)
  xt = torch.randint(0, V, (rows,), device=dev, generator=g)
  kw = dict(xt=xt, mc_t=0.7, mc_s=0.699, vocab=V, seq_len=L, round_bf16=True,
            fast_math=True, seed=11, step=0)
  # determinism
  a, _ = helpers.sampler_kernel(logits_a=la, logits_b=lb, diffusion='uniform',
                                guidance='cfg', gamma=2.0, **kw)
  b, _ = helpers.sampler_kernel(logits_a=la, logits_b=lb, diffusion='uniform',
                                guidance='cfg', gamma=2.0, **kw)
  assert torch.equal(a, b)
  assert int(a.min()) >= 0 and int(a.max()) < V
  # gamma == 1 mixes to the conditional posterior alone
  c, _ = helpers.sampler_kernel(logits_a=la, logits_b=lb, diffusion='uniform',
                                guidance='cfg', gamma=1.0, **kw)
  d, _ = helpers.sampler_kernel(logits_a=la, diffusion='uniform', guidance='none', **kw)
  assert (c != d).float().mean().item() < 2e-3
  # with mc_s ~ mc_t (tiny step) UDLM keeps almost every token
  e, _ = helpers.sampler_kernel(logits_a=la, diffusion='uniform', guidance='none',
                                **dict(kw, mc_t=0.5, mc_s=0.5 - 1e-6))
  assert (e == xt).float().mean().item() > 0.99
  # argmax selection == argmax of the analytic posterior (torch on the GPU)
  f, _ = helpers.sampler_kernel(logits_a=la, diffusion='uniform', guidance='none',
                                select='argmax', **dict(kw, round_bf16=False, fast_math=False))
  x = torch.softmax(la[:, :V].float(), dim=-1)
  a_s, a_t = 1 - 0.699, 1 - 0.7
  q = (a_s - a_t) * x + (1 - a_t / a_s) * (1 - a_s) / V
  q.scatter_add_(1, xt[:, None], (a_t * V * x.gather(1, xt[:, None]) + a_t / a_s - a_t))
  assert (f == q.argmax(-1)).float().mean().item() > 0.999
  # absorbing: unmasked tokens are carried over, masked ones never stay out of range
  xt_m = xt.clone()
  xt_m[::2] = 103
  h, _ = helpers.sampler_kernel(logits_a=la, diffusion='absorbing_state', guidance='none',
                                subs=True, mask_logit=True, mask_index=103,
                                **dict(kw, xt=xt_m))
  assert torch.equal(h[1::2], xt_m[1::2])


def test_ar_gumbel_selection():
  B, V = 16, 35
  dev = torch.device('cuda')
  g = torch.Generator().manual_seed(4)
  logits = torch.randn(B, 64, generator=g) * 2
  G = -torch.log(-torch.log(torch.rand(B, V, generator=g).clamp(1e-9, 1 - 1e-7)))
  lp = torch.log_softmax(logits[:, :V], -1)
  ref = ref_sampler.ar_select(lp, G)
  xs, _ = helpers.sampler_kernel(
    logits_a=logits.to(dev), xt=torch.zeros(B, dtype=torch.long), mc_t=0.0,
    mc_s=0.0, vocab=V, seq_len=1, diffusion='ar', noise=G, select='gumbel')
  assert torch.equal(xs.cpu(), ref)


def test_error_semantics():
  """Reference raises NotImplementedError for unknown diffusion / guidance
  (diffusion.py:1364-1366, 1408-1410)."""
  from ddg_b200 import _lib
  dev = torch.device('cuda')
  la = torch.zeros(4, 64, device=dev)
  xt = torch.zeros(4, dtype=torch.long, device=dev)
  a = _lib.SampleStepArgs(logits_a=la.data_ptr(), xt=xt.data_ptr(), xs=xt.data_ptr(),
                          rows=4, ld_a=64, seq_len=4, vocab=35, diffusion=7)
  with pytest.raises(NotImplementedError):
    _lib.check(_lib.lib.ddg_sample_step(a, 0))
  a.diffusion, a.guidance = 1, 9
  with pytest.raises(NotImplementedError):
    _lib.check(_lib.lib.ddg_sample_step(a, 0))


FUSED_MODES = [('uniform', 'none'), ('uniform', 'cfg'), ('absorbing_state', 'none')]


@pytest.mark.parametrize('round_bf16', [True, False])
@pytest.mark.parametrize('clf_round_bf16', [False, True])
@pytest.mark.parametrize('injected', [True, False])
def test_fused_single_pass_cbg_matches_generic_kernel(round_bf16, clf_round_bf16, injected):
  """UDLM D-CBG in the single-pass kernel (bf16 logits + GEMM partials + the fp32 classifier
  term in rows padded to whole 8-column vectors, as the first-order path produces it) against
  the careful generic kernel: ids equal except near-ties, with injected uniforms and with the
  kernels' own (identical) per-column Philox stream.  An unpadded term keeps the generic kernel."""
  B, L, V, ld = 4, 16, 30522, 30528
  rows = B * L
  dev = torch.device('cuda')
  la = _rand_logits(rows, V, ld, torch.bfloat16, 31).to(dev)
  g = torch.Generator().manual_seed(33)
  U = torch.rand(B, L, V, generator=g)
  xt = torch.randint(0, V, (B, L), generator=g)
  clf = torch.full((rows, ld), float('nan'))
  clf[:, :V] = torch.randn(rows, V, generator=g) * 1.5
  st_a = helpers.make_stats(la, V, -1)
  mc_t, mc_s = torch.tensor([0.61, 0.37, 0.9, 0.2]), torch.tensor([0.59, 0.33, 0.8, 0.1])
  common = dict(logits_a=la, clf=clf, noise=U if injected else None, xt=xt, mc_t=mc_t, mc_s=mc_s,
                vocab=V, seq_len=L, diffusion='uniform', guidance='cbg', gamma=2.0,
                round_bf16=round_bf16, clf_round_bf16=clf_round_bf16, seed=99, step=5)
  xs_ref, q_ref = helpers.sampler_kernel(fast_math=False, want_q=True, **common)
  fused = dict(fast_math=True, stats_a=st_a.data_ptr(), stats_ld=rows, stats_chunks=st_a.shape[0],
               race_group=64)
  xs, _ = helpers.sampler_kernel(**fused, **common)
  if injected:
    noise = U
  else:   # both kernels draw U_j from the same per-column counter (Philox streams)
    noise = torch.from_numpy(helpers.philox_uniform(99, 5, rows, V)).view(B, L, V)
  helpers.assert_ids_match(xs.view(B, L), xs_ref.view(B, L).cpu(), q_ref.view(B, L, V).cpu(), noise,
                           'fused uniform/cbg', tol=2e-2 if (round_bf16 or clf_round_bf16) else 1e-3)
  # a term without the row padding cannot be read as vectors: the generic kernel answers (same ids as
  # the generic reference run in its fast-math mode)
  xs_g1, _ = helpers.sampler_kernel(**fused, **dict(common, clf=clf[:, :V].contiguous()))
  xs_g2, _ = helpers.sampler_kernel(fast_math=True, **dict(common, clf=clf[:, :V].contiguous()))
  assert xs_g1.shape == xs_g2.shape


@pytest.mark.parametrize('diffusion,guidance', FUSED_MODES)
@pytest.mark.parametrize('round_bf16', [True, False])
def test_fused_single_pass_matches_generic_kernel(diffusion, guidance, round_bf16):
  """The production single-pass kernel (GEMM-supplied softmax partials, base-2
  intrinsics) against the careful generic kernel on the same logits and
  injected uniforms: ids equal except near-ties (|score gap| <= 1e-3 rel. with
  the intrinsics; 2e-2 when a bf16 rounding may flip)."""
  B, L, V, ld = 4, 16, 30522, 30528
  rows = B * L
  dev = torch.device('cuda')
  subs = diffusion == 'absorbing_state'
  mask = 103
  la = _rand_logits(rows, V, ld, torch.bfloat16, 21).to(dev)
  lb = _rand_logits(rows, V, ld, torch.bfloat16, 22).to(dev)
  g = torch.Generator().manual_seed(23)
  U = torch.rand(B, L, V, generator=g)
  if subs:
    xt = torch.full((B, L), mask, dtype=torch.long)
    xt[1, :3] = torch.tensor([5, 6, 7])
  else:
    xt = torch.randint(0, V, (B, L), generator=g)
  st_a = helpers.make_stats(la, V, mask if subs else -1)
  st_b = helpers.make_stats(lb, V, -1)
  mc_t, mc_s = torch.tensor([0.61, 0.37, 0.9, 0.2]), torch.tensor([0.59, 0.33, 0.8, 0.1])
  common = dict(logits_a=la, logits_b=lb if guidance == 'cfg' else None, noise=U,
                xt=xt, mc_t=mc_t, mc_s=mc_s, vocab=V, seq_len=L, diffusion=diffusion,
                guidance=guidance, gamma=2.0, subs=subs, mask_logit=subs,
                mask_index=mask, round_bf16=round_bf16)
  xs_ref, q_ref = helpers.sampler_kernel(fast_math=False, want_q=True, **common)
  xs, _ = helpers.sampler_kernel(
    fast_math=True, stats_a=st_a.data_ptr(),
    stats_b=st_b.data_ptr() if guidance == 'cfg' else 0, stats_ld=rows,
    stats_chunks=st_a.shape[0], **common)
  # make sure the fused kernel (not the generic one) answered: it ignores q_out
  helpers.assert_ids_match(xs.view(B, L), xs_ref.view(B, L).cpu(), q_ref.view(B, L, V).cpu(),
                           U, f'fused {diffusion}/{guidance}',
                           tol=2e-2 if round_bf16 else 1e-3)
  if subs:
    assert xs.view(B, L)[1, :3].tolist() == [5, 6, 7]


@pytest.mark.parametrize('diffusion,guidance', FUSED_MODES)
def test_fused_philox_matches_cpu_stream(diffusion, guidance):
  B, L, V, ld = 2, 8, 30522, 30528
  rows = B * L
  dev = torch.device('cuda')
  subs = diffusion == 'absorbing_state'
  la = _rand_logits(rows, V, ld, torch.bfloat16, 31).to(dev)
  lb = _rand_logits(rows, V, ld, torch.bfloat16, 32).to(dev)
  xt = (torch.full((B, L), 103, dtype=torch.long) if subs
        else torch.randint(0, V, (B, L), generator=torch.Generator().manual_seed(3)))
  st_a = helpers.make_stats(la, V, 103 if subs else -1)
  st_b = helpers.make_stats(lb, V, -1)
  seed, step = 0xFEDCBA9876543210, 901
  U = torch.from_numpy(helpers.philox_uniform(seed, step, rows, V))
  common = dict(logits_a=la, logits_b=lb if guidance == 'cfg' else None, xt=xt,
                mc_t=0.5, mc_s=0.05, vocab=V, seq_len=L, diffusion=diffusion,
                guidance=guidance, gamma=2.0, subs=subs, mask_logit=subs,
                mask_index=103, round_bf16=True, fast_math=True,
                stats_a=st_a.data_ptr(), stats_b=st_b.data_ptr() if guidance == 'cfg' else 0,
                stats_ld=rows, stats_chunks=st_a.shape[0])
  xs_inj, _ = helpers.sampler_kernel(noise=U, **common)
  xs_phx, _ = helpers.sampler_kernel(noise=None, seed=seed, step=step, **common)
  assert torch.equal(xs_inj, xs_phx)


def test_fused_sampler_statistics():
  """Philox production path draws from the right distribution: empirical
  token frequencies over many rows with identical logits match q_xs
  (chi-square style bound on the 50 most likely tokens)."""
  V, ld, n = 4096, 4096, 16384
  dev = torch.device('cuda')
  g = torch.Generator().manual_seed(5)
  row = (torch.randn(V, generator=g) * 2.5).to(torch.bfloat16)
  la = row[None].repeat(n, 1).contiguous().to(dev)
  xt = torch.full((n,), 17, dtype=torch.long)
  st = helpers.make_stats(la, V)
  xs, _ = helpers.sampler_kernel(
    logits_a=la, xt=xt, mc_t=0.8, mc_s=0.3, vocab=V, seq_len=1, diffusion='uniform',
    round_bf16=False, fast_math=True, seed=77, step=0, stats_a=st.data_ptr(),
    stats_ld=n, stats_chunks=st.shape[0])
  x = torch.softmax(row.float(), -1)
  from oracle import ref_sampler
  q = ref_sampler.posterior_uniform(x[None, None], torch.tensor([[17]]),
                                    torch.tensor(0.7), torch.tensor(0.2))[0, 0]
  counts = torch.bincount(xs.cpu(), minlength=V).float()
  top = q.topk(50).indices
  exp = q[top] * n
  zscore = (counts[top] - exp) / exp.clamp_min(1).sqrt()
  assert zscore.abs().max().item() < 5.0, zscore
  assert abs(counts[top].sum().item() - exp.sum().item()) < 5 * exp.sum().sqrt().item()


# --------------------------------------------------------------------------
# hierarchical race (production Philox path, csrc/sample_group.cu)
# --------------------------------------------------------------------------
def _group_inputs(diffusion, guidance, B, L, V, ld, seed0, scale=3.0):
  rows = B * L
  subs = diffusion == 'absorbing_state'
  la = _rand_logits(rows, V, ld, torch.bfloat16, seed0, scale)
  lb = _rand_logits(rows, V, ld, torch.bfloat16, seed0 + 1, scale)
  g = torch.Generator().manual_seed(seed0 + 2)
  if subs:
    xt = torch.full((B, L), 103, dtype=torch.long)
    xt[0, :2] = torch.tensor([7, 9])
  else:
    xt = torch.randint(0, V, (B, L), generator=g)
    xt[0, 0], xt[0, 1] = V - 1, V - 3          # z inside the ragged tail group
  return la, lb, xt, subs


def _oracle_probs(la, lb, xt, mc_t, mc_s, V, diffusion, guidance, gamma, round_bf16):
  B, L = xt.shape
  subs = diffusion == 'absorbing_state'
  lp_dtype = torch.bfloat16 if round_bf16 else None

  def lps(l):
    return ref_sampler.log_probs_from_logits(
      l[:, :V].float().view(B, L, V), torch.full_like(xt, 103) if subs else xt,
      parameterization='subs' if subs else 'd3pm', mask_index=103, lp_dtype=lp_dtype)
  U = torch.full((B, L, V), 0.5)
  kw = dict(diffusion=diffusion, mask_index=103, lp_dtype=lp_dtype)
  mct3, mcs3 = mc_t.view(-1, 1, 1), mc_s.view(-1, 1, 1)
  xt_all_mask = torch.full_like(xt, 103) if subs else xt   # posterior of the masked rows
  if guidance == 'cfg':
    _, q = ref_sampler.cfg_update(lps(la), lps(lb), gamma, xt_all_mask, mct3, mcs3, U, **kw)
  else:
    _, q = ref_sampler.ddpm_update(lps(la), xt_all_mask, mct3, mcs3, U, **kw)
  return q


@pytest.mark.parametrize('diffusion,guidance', FUSED_MODES)
@pytest.mark.parametrize('round_bf16', [True, False])
@pytest.mark.parametrize('race_group,V', [(8, 30522), (64, 30522), (64, 2051), (64, 4100),
                                          (64, 33001), (8, 33001)])
def test_group_race_matches_cpu_restatement(diffusion, guidance, round_bf16, race_group, V):
  """Production kernel (hierarchical race, in-kernel Philox) against the CPU
  restatement fed with the oracle's posterior and the regenerated Philox
  stream: ids bit-equal except near-ties -- a different group only if the two
  group scores agree to `tol` (log2 units), a different column of the same
  group only if the inverse-CDF target sits within `tol` of the boundary.
  Vocabulary sizes: the benchmark's, a ragged one just above the kernel's minimum, one
  whose last 64-column sweep is nearly empty, and one past 32768 (more than four sweeps
  per thread -> a second Philox call; softmax partials read from global memory)."""
  B, L = 4, 16
  ld = (V + 7) // 8 * 8
  rows = B * L
  dev = torch.device('cuda')
  la, lb, xt, subs = _group_inputs(diffusion, guidance, B, L, V, ld, 41)
  mc_t = torch.tensor([0.61, 0.37, 0.9, 0.2])
  mc_s = torch.tensor([0.59, 0.33, 0.8, 0.1])
  st_a = helpers.make_stats(la.to(dev), V, 103 if subs else -1)
  st_b = helpers.make_stats(lb.to(dev), V, -1)
  seed, step = 0x1234567ABCDEF01, 77
  xs, _ = helpers.sampler_kernel(
    logits_a=la.to(dev), logits_b=lb.to(dev) if guidance == 'cfg' else None, xt=xt,
    mc_t=mc_t, mc_s=mc_s, vocab=V, seq_len=L, diffusion=diffusion, guidance=guidance,
    gamma=2.0, subs=subs, mask_logit=subs, mask_index=103, round_bf16=round_bf16,
    fast_math=True, seed=seed, step=step, stats_a=st_a.data_ptr(),
    stats_b=st_b.data_ptr() if guidance == 'cfg' else 0, stats_ld=rows,
    stats_chunks=st_a.shape[0], race_group=race_group)
  xs = xs.cpu()
  q = _oracle_probs(la, lb, xt, mc_t, mc_s, V, diffusion, guidance, 2.0, round_bf16)
  ref = ref_sampler.group_race_select(q.view(rows, V), seed, step, group=race_group)
  ids_ref = ref['ids']
  xt_flat = xt.view(-1)
  if subs:     # carry-over rows keep their token (diffusion.py:1414-1418)
    keep = xt_flat != 103
    assert torch.equal(xs[keep], xt_flat[keep])
    ids_ref = torch.where(keep, xt_flat, ids_ref)
  tol = 3e-2 if round_bf16 else 2e-3
  bad = (xs != ids_ref).nonzero().flatten().tolist()
  def group_of(j):      # group index of column j
    if race_group == 8:
      return j // 8
    v = j // 8
    return ((v // 128) // 8) * 128 + v % 128
  for r in bad:
    gg, gr = group_of(int(xs[r])), group_of(int(ids_ref[r]))
    if gg != gr:
      gap = abs(ref['score'][r, gg] - ref['score'][r, gr])
      assert gap <= tol, f'row {r}: groups {gg} vs {gr}, score gap {gap:.3e}'
    else:
      cum, tgt = ref['cum'][r], ref['target'][r]
      cols = ref['cols'][r].tolist()
      lo, hi = sorted((cols.index(int(xs[r])), cols.index(int(ids_ref[r]))))
      edge = cum[lo:hi]
      assert (abs(edge - tgt) <= tol * cum[-1]).all(), f'row {r}: in-group pick not a near-tie'
  assert len(bad) <= max(2, rows // 16), f'{len(bad)} near-ties of {rows}'


@pytest.mark.parametrize('diffusion,guidance,round_bf16', [
  ('uniform', 'none', False), ('uniform', 'cfg', True), ('uniform', 'cfg', False),
  ('absorbing_state', 'none', True)])
def test_group_race_statistics(diffusion, guidance, round_bf16):
  """The hierarchical race draws from the reference's distribution: token
  frequencies over many rows with identical logits match the oracle posterior
  (z-scores of the 60 likeliest tokens and of the lumped remainder), and agree
  with the per-column race kernel on the same logits."""
  V, ld, n = 4096, 4096, 32768
  dev = torch.device('cuda')
  g = torch.Generator().manual_seed(11)
  ra = (torch.randn(V, generator=g) * 2.5).to(torch.bfloat16)
  rb = (ra.float() + torch.randn(V, generator=g) * 0.7).to(torch.bfloat16)
  la = ra[None].repeat(n, 1).contiguous().to(dev)
  lb = rb[None].repeat(n, 1).contiguous().to(dev)
  subs = diffusion == 'absorbing_state'
  z = 103 if subs else 17
  xt = torch.full((n,), z, dtype=torch.long)
  st_a = helpers.make_stats(la, V, 103 if subs else -1)
  st_b = helpers.make_stats(lb, V, -1)
  common = dict(
    logits_a=la, logits_b=lb if guidance == 'cfg' else None, xt=xt, mc_t=0.8, mc_s=0.3,
    vocab=V, seq_len=1, diffusion=diffusion, guidance=guidance, gamma=2.0, subs=subs,
    mask_logit=subs, mask_index=103, round_bf16=round_bf16, fast_math=True, seed=991,
    step=3, stats_a=st_a.data_ptr(), stats_b=st_b.data_ptr() if guidance == 'cfg' else 0,
    stats_ld=n, stats_chunks=st_a.shape[0])
  xs_grp, _ = helpers.sampler_kernel(race_group=8, **common)
  xs_wide, _ = helpers.sampler_kernel(race_group=64, **common)
  xs_col, _ = helpers.sampler_kernel(race_group=0, **common)
  q = _oracle_probs(ra[None], rb[None], xt[:1, None], torch.tensor([0.8]), torch.tensor([0.3]),
                    V, diffusion, guidance, 2.0, round_bf16)[0, 0].double()
  q = q / q.sum()
  top = q.topk(60).indices
  for xs, what in ((xs_grp, 'group'), (xs_wide, 'wide group'), (xs_col, 'column')):
    counts = torch.bincount(xs.cpu(), minlength=V).double()
    exp = q[top] * n
    zs = (counts[top] - exp) / (exp * (1 - q[top])).clamp_min(1).sqrt()
    assert zs.abs().max().item() < 5.0, (what, zs)
    rest, rest_exp = n - counts[top].sum().item(), n * (1 - q[top].sum().item())
    assert abs(rest - rest_exp) < 5 * max(rest_exp, 1) ** 0.5, (what, rest, rest_exp)


@pytest.mark.parametrize('race_group', [8, 64])
def test_group_race_full_size_properties(race_group):
  """Config-3 sized launch (64 x 128 rows of 30522) through the production
  kernel: ids in range, absorbing carry-over exact, deterministic for a fixed
  Philox key and different for another key."""
  rows, V, ld = 8192, 30522, 30528
  dev = torch.device('cuda')
  g = torch.Generator(device=dev).manual_seed(5)
  la = (torch.randn(rows, ld, device=dev, generator=g) * 2).to(torch.bfloat16)
  lb = (torch.randn(rows, ld, device=dev, generator=g) * 2).to(torch.bfloat16)
  xt = torch.randint(0, V, (rows,), device=dev, generator=g)
  st_a, st_b = helpers.make_stats(la, V), helpers.make_stats(lb, V)
  common = dict(logits_a=la, logits_b=lb, xt=xt, mc_t=0.5, mc_s=0.499, vocab=V, seq_len=128,
                diffusion='uniform', guidance='cfg', gamma=2.0, round_bf16=True, fast_math=True,
                stats_a=st_a.data_ptr(), stats_b=st_b.data_ptr(), stats_ld=rows,
                stats_chunks=st_a.shape[0], race_group=race_group)
  a1, _ = helpers.sampler_kernel(seed=1, step=5, **common)
  a2, _ = helpers.sampler_kernel(seed=1, step=5, **common)
  a3, _ = helpers.sampler_kernel(seed=2, step=5, **common)
  assert torch.equal(a1, a2)
  assert int(a1.min()) >= 0 and int(a1.max()) < V
  # with dt = 1/1000 most tokens stay (UDLM posterior mass at z): sanity of the z term
  assert (a1 == xt).float().mean().item() > 0.9
  assert (a1 != a3).any()


@pytest.mark.parametrize('diffusion', ['uniform', 'absorbing_state'])
@pytest.mark.parametrize('round_bf16', [True, False])
def test_float64_kernel_vs_oracle(diffusion, round_bf16):
  """sampling.use_float64 (diffusion.py:1393-1394): log-probs at the forward's
  precision (bf16 when round_bf16), everything after them in double."""
  B, L, V, ld = 2, 6, 1000, 1000
  rows = B * L
  dev = torch.device('cuda')
  subs = diffusion == 'absorbing_state'
  mask = 103
  la = _rand_logits(rows, V, ld, torch.bfloat16, 5)
  g = torch.Generator().manual_seed(6)
  U = torch.rand(B, L, V, generator=g)
  xt = (torch.full((B, L), mask, dtype=torch.long) if subs
        else torch.randint(0, V, (B, L), generator=g))
  if subs:
    xt[1, :2] = torch.tensor([3, 8])
  mc_t, mc_s = torch.tensor([0.61, 0.37]), torch.tensor([0.59, 0.33])
  lp_dtype = torch.bfloat16 if round_bf16 else None
  lp = ref_sampler.log_probs_from_logits(
    la[:, :V].float().view(B, L, V), xt, parameterization='subs' if subs else 'd3pm',
    mask_index=mask, lp_dtype=lp_dtype)
  xs_ref, q_ref = ref_sampler.ddpm_update(
    lp, xt, mc_t[:, None, None], mc_s[:, None, None], U, diffusion=diffusion, mask_index=mask,
    lp_dtype=lp_dtype, use_float64=True)
  assert q_ref.dtype == torch.float64
  xs, q = helpers.sampler_kernel(
    logits_a=la.to(dev), noise=U, xt=xt, mc_t=mc_t, mc_s=mc_s, vocab=V, seq_len=L,
    diffusion=diffusion, subs=subs, mask_logit=subs, mask_index=mask, round_bf16=round_bf16,
    want_q=True, compute_f64=True)
  helpers.assert_close_frac(q.view(B, L, V), q_ref, rtol=3e-4, atol=1e-7,
                            loose_frac=1e-3 if round_bf16 else 0.0, loose_rtol=2e-2)
  helpers.assert_ids_match(xs.view(B, L), xs_ref, q_ref.float(), U, f'f64 {diffusion}', tol=2e-5)


@pytest.mark.parametrize('diffusion,guidance', [('uniform', 'cbg'), ('absorbing_state', 'cbg'),
                                                ('absorbing_state', 'cfg')])
def test_generic_kernel_with_gemm_partials_equals_two_pass(diffusion, guidance):
  """Guided modes served by the generic kernel (D-CBG, MDLM D-CFG) in throughput mode
  (Philox noise, fast math): with the vocab GEMM's log-softmax partials the normaliser
  pass over the row is skipped; posterior and ids must equal the two-pass run."""
  B, L, V, ld = 3, 8, 30522, 30528
  rows = B * L
  dev = torch.device('cuda')
  subs = diffusion == 'absorbing_state'
  la = _rand_logits(rows, V, ld, torch.bfloat16, 21, 2.0).to(dev)
  lb = _rand_logits(rows, V, ld, torch.bfloat16, 22, 2.0).to(dev)
  g = torch.Generator().manual_seed(23)
  clf = torch.randn(B, L, V, generator=g) * 0.5 - 1.0
  if subs:
    xt = torch.full((B, L), 103, dtype=torch.long)
    xt[1, :3] = torch.tensor([5, 6, 7])
  else:
    xt = torch.randint(0, V, (B, L), generator=g)
  st_a = helpers.make_stats(la, V, 103 if subs else -1)
  st_b = helpers.make_stats(lb, V, 103 if subs else -1)
  common = dict(logits_a=la, logits_b=lb if guidance == 'cfg' else None,
                clf=clf if guidance == 'cbg' else None, xt=xt, mc_t=0.6, mc_s=0.55, vocab=V,
                seq_len=L, diffusion=diffusion, guidance=guidance, gamma=1.5, subs=subs,
                mask_logit=subs, mask_index=103, round_bf16=True, fast_math=True, seed=77, step=2,
                want_q=True, clf_round_bf16=guidance == 'cbg')
  xs0, q0 = helpers.sampler_kernel(**common)
  xs1, q1 = helpers.sampler_kernel(stats_a=st_a.data_ptr(),
                                   stats_b=st_b.data_ptr() if guidance == 'cfg' else 0,
                                   stats_ld=rows, stats_chunks=st_a.shape[0], **common)
  helpers.assert_close_frac(q1.cpu(), q0.cpu(), rtol=3e-4, atol=1e-7, loose_frac=2e-3, loose_rtol=2e-2)
  assert (xs0 != xs1).float().mean().item() <= 0.05
  if subs:
    assert xs1.view(B, L)[1, :3].tolist() == [5, 6, 7]
