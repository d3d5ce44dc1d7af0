"""Pins the CPU oracle (oracle/) to the UNMODIFIED reference: every golden
fixture under tests/golden/ was produced by the reference sampler itself
(make_golden.py); the oracle must reproduce ids exactly and log-probs /
posteriors to fp32 round-off on the same seeded inputs."""
import pytest
import torch

import helpers
from golden import cases
from oracle import ref_dit, ref_sampler

DIFFUSION_CASES = [c.name for c in cases.CASES if c.parameterization != 'ar' and not c.use_cache]


@pytest.mark.parametrize('name', DIFFUSION_CASES)
def test_oracle_replays_reference(name):
  case = cases.CASE_BY_NAME[name]
  gold = helpers.load_golden(name)
  big = case.vocab == 'bert'
  oracle = helpers.make_oracle(case)
  B = case.batch
  sig, mct, mcs = ref_sampler.step_scalars(case.steps, 1e-5, case.T)
  cond = None
  if case.guidance == 'cfg':
    cond = torch.full((B,), case.condition, dtype=torch.long)
  xt = torch.from_numpy(gold['xt0']).long()
  for i in range(case.steps):
    assert torch.equal(xt, torch.from_numpy(gold[f's{i}_xt']).long())
    sc = gold[f's{i}_scalars']
    assert abs(sig[i].item() - sc[0]) <= 1e-6 * max(1, abs(sc[0]))
    assert abs(mct[i].item() - sc[1]) <= 1e-6 and abs(mcs[i].item() - sc[2]) <= 1e-6
    sigma_t = sig[i].expand(B)
    mc_t = mct[i].expand(B)[:, None, None]
    mc_s = mcs[i].expand(B)[:, None, None]
    xs, q, lps = oracle.step(
      xt, sigma_t, mc_t, mc_s, cases.step_noise(case, i),
      guidance=case.guidance, cond=cond, gamma=case.gamma,
      condition=case.condition, use_approx=case.use_approx)
    for k, v in lps.items():
      if v is not None:
        # -1e6 entries of the subs parameterization make rtol the right gauge
        helpers.check_summary(f's{i}_{k}', v, gold, big, rtol=2e-4, atol=2e-4)
    helpers.check_summary(f's{i}_q_xs', q, gold, big, rtol=2e-4, atol=1e-7)
    gold_xs = torch.from_numpy(gold[f's{i}_xs']).long()
    assert torch.equal(xs, gold_xs), (
      f'step {i}: {(xs != gold_xs).sum().item()} ids differ')
    xt = gold_xs
  assert torch.equal(xt, torch.from_numpy(gold['samples']).long())


@pytest.mark.parametrize('name', ['ar_null', 'ar_cfg_g2', 'ar_lowconf'])
def test_oracle_replays_reference_ar(name):
  """AR sampler incl. `_ar_cfg_denoise` (diffusion.py:1188-1236) and the
  low-confidence `_sample_token` (:1024-1038): ids bit-exact."""
  case = cases.CASE_BY_NAME[name]
  gold = helpers.load_golden(name)
  cfg, tok, sd, _ = helpers.case_state(case)
  G = cases.ar_gumbel(case)
  B = case.batch
  x = torch.zeros(B, case.length, dtype=torch.long)
  x[:, 0] = tok.bos_token_id
  cond = torch.full((B,), case.condition, dtype=torch.long)
  uncond = torch.full((B,), case.num_classes, dtype=torch.long)
  lowc = case.low_confidence_threshold if case.low_confidence else None

  def log_probs(ids, c):
    logits = ref_dit.dit_forward(sd, ids, None, c, n_heads=case.n_heads, causal=True)
    return ref_sampler.log_probs_from_logits(logits, ids, parameterization='ar',
                                             mask_index=tok.mask_token_id)
  for i in range(case.length - 1):
    if case.guidance == 'cfg':
      lp = ref_sampler.ar_cfg_mix(log_probs(x[:, :i + 1], cond)[:, -1],
                                  log_probs(x[:, :i + 1], uncond)[:, -1], case.gamma)
    else:
      lp = log_probs(x[:, :i + 1], None)[:, -1]
    x[:, i + 1] = ref_sampler.ar_select(lp, G[:, i], low_confidence=lowc)
  assert torch.equal(x, torch.from_numpy(gold['samples']).long())
  # causal: the full-length forward reproduces every prefix forward's last row
  if case.guidance == 'cfg':
    helpers.check_summary('log_probs_final_cond', log_probs(x, cond), gold, False, 2e-4, 2e-4)
    helpers.check_summary('log_probs_final_uncond', log_probs(x, uncond), gold, False, 2e-4, 2e-4)
  else:
    helpers.check_summary('log_probs_final', log_probs(x, None), gold, False, 2e-4, 2e-4)


# ---------------------------------------------------------------- GPU rounding points
# `<case>_amp.npz`: the unmodified reference run with its torch.cuda.amp.autocast
# blocks mapped onto CPU autocast of the same dtype (make_golden.py --amp), i.e.
# rounded where the reference rounds on a GPU.  A bf16 value moves by >= 0.4 % when a
# rounding is added, dropped or moved, while two implementations that round at the
# same points agree BIT FOR BIT wherever the fp32 accumulation-order noise upstream
# did not flip a rounding.  So the pin is the FRACTION of bit-identical entries: it
# is 0.6 - 0.95 for the oracle's gpu_rounding mode and 0.00 for its fp32 mode.
EXPECTED_DTYPES = {'q_xs': 'torch.float32', 'log_x_theta': 'torch.bfloat16',
                   'log_x_theta_cond': 'torch.bfloat16',
                   'log_x_theta_uncond': 'torch.bfloat16',
                   # fp32 on a GPU as well: log_softmax under autocast(float32)
                   # (classifier.py:325-327); measured on the B200 (`*_gpu.npz`)
                   'classifier_log_prob': 'torch.float32'}


def _amp_replay(case, gold, gpu_rounding):
  """Yields (step, key, oracle tensor, golden tensor) plus ('xs', ...) per step."""
  big = case.vocab == 'bert'
  oracle = helpers.make_oracle(case, gpu_rounding=gpu_rounding)
  B = case.batch
  sig, mct, mcs = ref_sampler.step_scalars(case.steps, 1e-5, case.T)
  cond = (torch.full((B,), case.condition, dtype=torch.long)
          if case.guidance == 'cfg' else None)
  for i in range(case.steps):
    xt = torch.from_numpy(gold[f's{i}_xt']).long()
    xs, q, lps = oracle.step(
      xt, sig[i].expand(B), mct[i].expand(B)[:, None, None],
      mcs[i].expand(B)[:, None, None], cases.step_noise(case, i),
      guidance=case.guidance, cond=cond, gamma=case.gamma,
      condition=case.condition, use_approx=case.use_approx)
    yield i, 'xs', xs, torch.from_numpy(gold[f's{i}_xs']).long()
    for k, v in list(lps.items()) + [('q_xs', q)]:
      if v is None:
        continue
      g = torch.from_numpy(gold[f's{i}_{k}' + ('_cols' if big else '')]).float()
      yield i, k, (v.float()[..., ::997] if big else v.float()), g


@pytest.mark.parametrize('source', ['amp', 'gpu'])
@pytest.mark.parametrize('name', cases.AMP_CASES)
def test_oracle_rounding_points_match_reference_under_autocast(name, source):
  """`amp`: the reference with its cuda autocast policy emulated on the CPU
  (make_golden.py --amp); `gpu`: the reference's OWN run on a B200 under real cuda
  autocast -- cuBLAS / flash-attn / SDPA kernels -- recorded by
  `make_golden.py --gpu` (tests/golden/<case>_gpu.npz)."""
  case = cases.CASE_BY_NAME[name]
  gold = helpers.load_golden(f'{name}_{source}')
  want_dtypes = dict(EXPECTED_DTYPES)
  for k, want in want_dtypes.items():
    if f'{k}_dtype' in gold:
      assert str(gold[f'{k}_dtype']) == want, (k, gold[f'{k}_dtype'])
  n_ids = n_diff = 0
  for i, k, v, g in _amp_replay(case, gold, gpu_rounding=True):
    if k == 'xs':
      n_ids += v.numel()
      n_diff += int((v != g).sum())
      continue
    fin = torch.isfinite(g) & (g > -1e5) & (g != 0)
    if k == 'q_xs' or want_dtypes[k] == 'torch.float32':
      # every entry of a posterior row depends on the whole row's log-probs through
      # the normaliser, so no bit-level statistic here (the sampler stage is pinned
      # exactly by the next test); bf16-ulp-level agreement only
      # (same for the fp32 first-order classifier term: a bf16 backward upstream)
      rel = ((v - g).abs() / g.abs())[fin]
      assert rel.max().item() <= 0.3, (name, i, k, rel.max().item())
      assert rel.median().item() <= 3e-2, (name, i, k, rel.median().item())
    else:
      assert bool((v == v.bfloat16().float()).all()), (name, i, k, 'not bf16 values')
      rel = ((v - g).abs() / g.abs())[fin]
      same = (rel == 0).float().mean().item()
      assert same >= 0.25, (name, i, k, same)
      assert rel.max().item() <= 2.5e-2, (name, i, k, rel.max().item())   # <= 3 bf16 ulps
  assert n_diff <= max(2, n_ids // 20), (name, n_diff, n_ids)   # near-tie flips only


@pytest.mark.parametrize('name', ['udlm_cfg_g2', 'mdlm_null'])
def test_fp32_oracle_is_told_apart_by_the_rounding_pin(name):
  """The same statistic rejects an implementation that does not round: the fp32
  oracle has (almost) no log-prob entry in common with the rounded reference."""
  case = cases.CASE_BY_NAME[name]
  gold = helpers.load_golden(name + '_amp')
  for i, k, v, g in _amp_replay(case, gold, gpu_rounding=False):
    if k.startswith('log_x_theta') and i == 0:
      fin = torch.isfinite(g) & (g > -1e5) & (g != 0)
      assert ((v - g)[fin] == 0).float().mean().item() < 0.05, (name, k)


SMALL_AMP_CASES = [n for n in cases.AMP_CASES if cases.CASE_BY_NAME[n].vocab != 'bert']


@pytest.mark.parametrize('source', ['amp', 'gpu'])
@pytest.mark.parametrize('name', SMALL_AMP_CASES)
def test_sampler_stage_rounding_is_exact_on_reference_log_probs(name, source):
  """Feeds the reference's own bf16 log-probs (and fp32 classifier term) of every
  traced step into the oracle's per-step update with the GPU rounding switches on:
  with identical inputs, the posterior must agree to fp32 round-off and the ids
  exactly -- exp / D-CFG mix / softmax round where the reference rounds them
  (diffusion.py:1395, 1476-1477, 1484-1489, 1596).  `gpu`: the log-probs, posterior
  and ids are those of the reference's own B200 run."""
  case = cases.CASE_BY_NAME[name]
  gold = helpers.load_golden(f'{name}_{source}')
  tok = cases.Tok(case.vocab)
  B = case.batch
  kw = dict(diffusion=case.diffusion, mask_index=tok.mask_token_id,
            lp_dtype=torch.bfloat16)
  for i in range(case.steps):
    xt = torch.from_numpy(gold[f's{i}_xt']).long()
    sc = gold[f's{i}_scalars']
    mc_t = torch.full((B, 1, 1), float(sc[1]))
    mc_s = torch.full((B, 1, 1), float(sc[2]))
    U = cases.step_noise(case, i)
    lp = {k: torch.from_numpy(gold[f's{i}_{k}']).float()
          for k in ('log_x_theta', 'log_x_theta_cond', 'log_x_theta_uncond',
                    'classifier_log_prob') if f's{i}_{k}' in gold}
    if case.guidance is None:
      xs, q = ref_sampler.ddpm_update(lp['log_x_theta'], xt, mc_t, mc_s, U, **kw)
    elif case.guidance == 'cfg':
      xs, q = ref_sampler.cfg_update(lp['log_x_theta_cond'], lp['log_x_theta_uncond'],
                                     case.gamma, xt, mc_t, mc_s, U, **kw)
    else:
      # the classifier term is fp32 (classifier.py:325-327): gamma * term stays fp32
      xs, q = ref_sampler.cbg_update(lp['log_x_theta'], lp['classifier_log_prob'],
                                     case.gamma, xt, mc_t, mc_s, U, **kw)
    g = torch.from_numpy(gold[f's{i}_q_xs']).float()
    torch.testing.assert_close(q.float(), g, rtol=2e-5, atol=1e-9)
    assert torch.equal(xs, torch.from_numpy(gold[f's{i}_xs']).long()), (name, i)
    # and the switches matter: without the bf16 roundings the posterior moves by > 1e-4
    # (MDLM + D-CBG has no rounding after the log-probs: nothing to tell apart)
    if not (case.guidance == 'cbg' and case.diffusion == 'absorbing_state'):
      kw32 = dict(kw, lp_dtype=None)
      if case.guidance is None:
        _, q32 = ref_sampler.ddpm_update(lp['log_x_theta'], xt, mc_t, mc_s, U, **kw32)
      elif case.guidance == 'cfg':
        _, q32 = ref_sampler.cfg_update(lp['log_x_theta_cond'], lp['log_x_theta_uncond'],
                                        case.gamma, xt, mc_t, mc_s, U, **kw32)
      else:
        _, q32 = ref_sampler.cbg_update(lp['log_x_theta'], lp['classifier_log_prob'],
                                        case.gamma, xt, mc_t, mc_s, U, **kw32)
      pos = g > 1e-12
      assert ((q32.float() - g).abs() / g.clamp_min(1e-30))[pos].max().item() > 1e-4


@pytest.mark.parametrize('name', ['udlm_use_cache', 'udlm_cfg_use_cache'])
def test_oracle_replays_reference_use_cache_rule(name):
  """`sampling.use_cache` with time conditioning (diffusion.py:1371-1374): while no
  token changes, the reference re-uses the log-probs computed at an OLDER sigma.
  The traced `cached` flags, ids and posteriors of the reference must follow from:
  keep the cache iff xs == xt."""
  case = cases.CASE_BY_NAME[name]
  gold = helpers.load_golden(name)
  oracle = helpers.make_oracle(case)
  tok = cases.Tok(case.vocab)
  B = case.batch
  sig, mct, mcs = ref_sampler.step_scalars(case.steps, 1e-5, case.T)
  cond = torch.full((B,), case.condition, dtype=torch.long) if case.guidance == 'cfg' else None
  kw = dict(diffusion=case.diffusion, mask_index=tok.mask_token_id)
  xt = torch.from_numpy(gold['xt0']).long()
  cache, n_cached = None, 0
  for i in range(case.steps):
    assert torch.equal(xt, torch.from_numpy(gold[f's{i}_xt']).long())
    assert int(gold[f's{i}_cached']) == int(cache is not None), (name, i)
    mc_t, mc_s = mct[i].expand(B)[:, None, None], mcs[i].expand(B)[:, None, None]
    U = cases.step_noise(case, i)
    if cache is None:
      xs, q, lps = oracle.step(xt, sig[i].expand(B), mc_t, mc_s, U, guidance=case.guidance,
                               cond=cond, gamma=case.gamma)
    else:
      n_cached += 1
      lps = cache
      if case.guidance == 'cfg':
        xs, q = ref_sampler.cfg_update(lps['log_x_theta_cond'], lps['log_x_theta_uncond'],
                                       case.gamma, xt, mc_t, mc_s, U, **kw)
      else:
        xs, q = ref_sampler.ddpm_update(lps['log_x_theta'], xt, mc_t, mc_s, U, **kw)
    assert torch.equal(xs, torch.from_numpy(gold[f's{i}_xs']).long()), (name, i)
    helpers.check_summary(f's{i}_q_xs', q, gold, False, 2e-4, 1e-6)
    cache = lps if torch.equal(xs, xt) else None
    xt = xs
  assert n_cached >= 3
