"""Pins the CPU oracle (oracle/) to the UNMODIFIED reference: every golden
fixture under tests/golden/ was produced by the reference sampler itself
(make_golden.py); the oracle must reproduce ids exactly and log-probs /
posteriors to fp32 round-off on the same seeded inputs."""
import pytest
import torch

import helpers
from golden import cases
from oracle import ref_dit, ref_sampler

DIFFUSION_CASES = [c.name for c in cases.CASES if c.parameterization != 'ar']


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


def test_oracle_replays_reference_ar():
  case = cases.CASE_BY_NAME['ar_null']
  gold = helpers.load_golden('ar_null')
  cfg, tok, sd, _ = helpers.case_state(case)
  G = cases.ar_gumbel(case)
  x = torch.zeros(case.batch, case.length, dtype=torch.long)
  x[:, 0] = tok.bos_token_id
  for i in range(case.length - 1):
    logits = ref_dit.dit_forward(sd, x[:, :i + 1], None, None,
                                 n_heads=case.n_heads, causal=True)
    lp = ref_sampler.log_probs_from_logits(
      logits, x[:, :i + 1], parameterization='ar',
      mask_index=tok.mask_token_id)
    x[:, i + 1] = ref_sampler.ar_select(lp[:, -1], G[:, i])
  assert torch.equal(x, torch.from_numpy(gold['samples']).long())
  # causal: the full-length forward reproduces every prefix forward's last row
  logits = ref_dit.dit_forward(sd, x, None, None, n_heads=case.n_heads,
                               causal=True)
  lp = ref_sampler.log_probs_from_logits(
    logits, x, parameterization='ar', mask_index=tok.mask_token_id)
  helpers.check_summary('log_probs_final', lp, gold, False, 2e-4, 2e-4)
