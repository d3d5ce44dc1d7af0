"""Oracle-compared parity at the BENCHMARKED widths and of the entry points that
had no oracle test: the whole DiT-small forward at
hidden 768 x 12 heads x 12 blocks (C2 / C3 / C5) and DiT-tiny at L = 256 (C1)
against `oracle.ref_dit`, `ddg_log_probs` against
`ref_sampler.log_probs_from_logits`, `ddg_sample_prior` against its regenerated
Philox stream, and the AR D-CFG step against `ar_cfg_mix` + `ar_select`."""
import numpy as np
import pytest
import torch

import ddg_b200
import helpers
from ddg_b200 import _lib
from golden import cases
from oracle import ref_dit, ref_sampler

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _forward_vs_oracle(case, seed, cond_classes=0):
  cfg, tok, sd, _ = helpers.case_state(case)
  model = ddg_b200.Diffusion(cfg, tok, device=DEV).load_state_dict(sd)
  B, L, V = case.batch, case.length, tok.vocab_size
  g = torch.Generator().manual_seed(seed)
  ids = torch.randint(0, V, (B, L), generator=g)
  sigma = torch.rand(B, generator=g) * 4
  cond = torch.randint(0, cond_classes, (B,), generator=g) if cond_classes else None
  logits = model._logits(ids.to(DEV), sigma.to(DEV), None if cond is None else cond.to(DEV))
  torch.cuda.synchronize()
  with torch.no_grad():
    ref = ref_dit.dit_forward(sd, ids, sigma, cond, n_heads=case.n_heads,
                              rounding=ref_dit.Rounding.gpu_denoiser())
    ref32 = ref_dit.dit_forward(sd, ids, sigma, cond, n_heads=case.n_heads)
  got = logits.float().cpu()
  d = (got - ref).abs()
  print(f'[width] {case.name}: rows={B * L} max|d|={d.max().item():.4f} '
        f'mean|d|={d.mean().item():.5f} vs fp32 mean|d|={(got - ref32).abs().mean().item():.5f} '
        f'bit-equal {(got == ref).float().mean().item():.3f}')
  # same bound as tests/test_gpu_dit.py::test_dit_forward_matches_oracle (bf16 logits:
  # 1 ulp = 2^-8 relative)
  assert bool((d <= 0.06 + 0.03 * ref.abs()).all()), d.max().item()
  assert d.mean().item() <= 0.01, d.mean().item()
  assert (got - ref32).abs().mean().item() <= 0.03
  return model, got, ref, ids, sigma, cond


def test_dit_small_forward_at_benchmark_width_matches_oracle():
  """C2 / C3 / C5's network: hidden 768, 12 heads, **12 blocks**, L = 128,
  V = 30522, 32 sequences = 4096 token rows -> streaming LayerNorm at hidden 768,
  tcgen05 attention at 12 heads, 2-CTA 256x256 GEMM tiles, the padded vocab head
  with its log-softmax partials; 12 blocks of error growth."""
  case = cases.Case('small768', vocab='bert', guidance='cfg', gamma=2.0, batch=32, length=128,
                    hidden=768, n_heads=12, n_blocks=12, cond_dim=128, seed=5)
  model, got, ref, ids, sigma, cond = _forward_vs_oracle(case, 17, cond_classes=3)
  # the vocab GEMM's log-softmax partials at this width: log-probs from (logits, partials)
  # through the C ABI == log_softmax of the same logits
  lp = model.forward(ids.to(DEV), sigma.to(DEV), cond.to(DEV)).cpu()
  exp = ref_sampler.log_probs_from_logits(got, ids, parameterization='d3pm',
                                          mask_index=103, lp_dtype=torch.bfloat16)
  assert (lp - exp).abs().max().item() <= 0.07    # one bf16 ulp at |lp| ~ 16


def test_dit_tiny_l256_forward_matches_oracle():
  """C1's network (text8 UDLM DiT-tiny: hidden 512, 8 heads, 8 blocks) at
  L = 256, V = 35: the whole forward through `attention_tc_kernel<., 256>` and,
  with 20 sequences = 5120 rows, the streaming LayerNorm at hidden 512."""
  case = cases.Case('tiny256', vocab='text8', batch=20, length=256, hidden=512, n_heads=8,
                    n_blocks=8, cond_dim=128, seed=6)
  _forward_vs_oracle(case, 19)


def test_dit_tiny_l256_batch4_matches_oracle():
  """C1 at its own batch of 4 (1024 rows: the generic LayerNorm kernel)."""
  case = cases.Case('tiny256b4', vocab='text8', batch=4, length=256, hidden=512, n_heads=8,
                    n_blocks=8, cond_dim=128, seed=7)
  _forward_vs_oracle(case, 23)


@pytest.mark.parametrize('param,diffusion,subs_masking', [
  ('subs', 'absorbing_state', False), ('d3pm', 'uniform', False),
  ('d3pm', 'absorbing_state', True), ('ar', 'absorbing_state', False)])
@pytest.mark.parametrize('vocab', ['text8', 'bert'])
def test_log_probs_entry_point_matches_oracle(param, diffusion, subs_masking, vocab):
  """`ddg_log_probs` (what `Diffusion.forward` returns, reference
  diffusion.py:350-365 incl. `_subs_parameterization` :321-337) against the
  oracle on the same bf16 logits, with and without the reference's bf16
  rounding of the log-softmax."""
  tok = cases.Tok(vocab)
  V, m = tok.vocab_size, tok.mask_token_id
  rows, ld = 96, (V + 63) // 64 * 64
  g = torch.Generator().manual_seed(31)
  logits = (torch.randn(rows, ld, generator=g) * 3).to(torch.bfloat16)
  xt = torch.randint(0, V, (rows,), generator=g)
  xt[::3] = m                                   # a third of the rows still masked
  for round_bf16 in (0, 1):
    out = torch.empty(rows, V, dtype=torch.float32, device=DEV)
    lg = logits.to(DEV)
    xd = xt.to(DEV)
    _lib.check(_lib.lib.ddg_log_probs(
      lg.data_ptr(), ld, _lib.BF16, xd.data_ptr(), rows, V,
      1 if param == 'subs' else 0, 1 if (param == 'subs' or subs_masking) else 0, m,
      round_bf16, out.data_ptr(), V, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = ref_sampler.log_probs_from_logits(
      logits[:, :V].float(), xt, parameterization=param, mask_index=m,
      subs_masking=subs_masking, lp_dtype=torch.bfloat16 if round_bf16 else None)
    got = out.cpu()
    big = ref < -1e5                            # -1e6-type entries: compare loosely
    assert bool((got[big] < -1e5).all())
    if round_bf16:   # both are bf16 values: equal except where fp32 sits on a rounding edge
      rel = ((got - ref).abs() / ref.abs().clamp_min(1e-6))[~big]
      assert rel.max().item() <= 2 ** -7, rel.max().item()
      assert (rel == 0).float().mean().item() >= 0.98
    else:
      torch.testing.assert_close(got[~big], ref[~big], rtol=2e-6, atol=2e-5)


def test_sample_prior_matches_regenerated_stream():
  """`ddg_sample_prior` (reference diffusion.py:1006-1013).  absorbing: all
  [MASK]; uniform: `randint(0, V)` drawn from Philox4x32-10 with counter
  (i/4, 0, 0, 0xFFFFFFFF), id = (word * V) >> 32 -- regenerated on the CPU bit
  for bit; in range; chi-square against the uniform law."""
  V, n = 30522, 512 * 128 + 3           # not a multiple of 4: tail handled
  for seed in (1, (7 << 32) | 12345):
    xt = torch.empty(n, dtype=torch.int64, device=DEV)
    _lib.check(_lib.lib.ddg_sample_prior(xt.data_ptr(), n, _lib.UNIFORM, V, 103, seed,
                                         _lib.stream_ptr()))
    torch.cuda.synchronize()
    n4 = (n + 3) // 4
    i4 = np.arange(n4, dtype=np.uint64)
    w = ref_sampler.philox4x32_10(
      (i4 & np.uint64(0xFFFFFFFF)).astype(np.uint32), (i4 >> np.uint64(32)).astype(np.uint32),
      np.zeros(n4, np.uint32), np.full(n4, 0xFFFFFFFF, np.uint32),
      seed & 0xFFFFFFFF, seed >> 32)
    words = np.stack(w, axis=-1).reshape(-1)[:n].astype(np.uint64)
    exp = ((words * np.uint64(V)) >> np.uint64(32)).astype(np.int64)
    got = xt.cpu().numpy()
    assert np.array_equal(got, exp)
    assert got.min() >= 0 and got.max() < V
  # chi-square over 64 equal-width bins of the id range (last seed's draw)
  counts = np.bincount(got * 64 // V, minlength=64).astype(np.float64)
  expected = np.diff(np.ceil(np.arange(65) * V / 64)).astype(np.float64) * n / V
  chi2 = ((counts - expected) ** 2 / expected).sum()
  assert chi2 < 120.0, chi2              # 63 dof: P(chi2 > 120) ~ 2e-5
  # two seeds give different draws
  xt2 = torch.empty(n, dtype=torch.int64, device=DEV)
  _lib.check(_lib.lib.ddg_sample_prior(xt2.data_ptr(), n, _lib.UNIFORM, V, 103, 1,
                                       _lib.stream_ptr()))
  assert (xt2.cpu().numpy() != got).mean() > 0.99
  xa = torch.empty(1000, dtype=torch.int64, device=DEV)
  _lib.check(_lib.lib.ddg_sample_prior(xa.data_ptr(), 1000, _lib.ABSORBING, V, 103, 5,
                                       _lib.stream_ptr()))
  assert bool((xa == 103).all())


def test_sample_prior_through_diffusion_api():
  """`Diffusion._sample_prior` for both diffusions (no prior_fn injected)."""
  for name in ('udlm_null', 'mdlm_null'):
    case = cases.CASE_BY_NAME[name]
    cfg, tok, sd, _ = helpers.case_state(case)
    model = ddg_b200.Diffusion(cfg, tok, device=DEV).load_state_dict(sd)
    xt = model._sample_prior(64, 32)
    assert xt.dtype == torch.int64 and tuple(xt.shape) == (64, 32)
    if case.diffusion == 'absorbing_state':
      assert bool((xt == tok.mask_token_id).all())
    else:
      assert int(xt.min()) >= 0 and int(xt.max()) < tok.vocab_size
      assert len(torch.unique(xt)) > tok.vocab_size // 2


@pytest.mark.parametrize('gamma', [2.0, 0.5])
@pytest.mark.parametrize('vocab', ['text8', 'bert'])
def test_ar_cfg_step_matches_oracle(gamma, vocab):
  """AR D-CFG token step (reference `_ar_cfg_denoise`, diffusion.py:1188-1236,
  then `_sample_token` :1024-1038): ids of the fused kernel on fp32 logits and
  injected Gumbel noise against `ar_select(ar_cfg_mix(lp_c, lp_u, gamma), G)`;
  differences only where the two candidates' scores tie to 2e-5."""
  V = cases.Tok(vocab).vocab_size
  B, ld = 64, (V + 63) // 64 * 64
  g = torch.Generator().manual_seed(41)
  la = torch.randn(B, ld, generator=g) * 2.5
  lb = la * 0.6 + torch.randn(B, ld, generator=g)
  G = -torch.log(-torch.log(torch.rand(B, V, generator=g).clamp(1e-9, 1 - 1e-7)))
  lp_c = torch.log_softmax(la[:, :V].double(), -1)
  lp_u = torch.log_softmax(lb[:, :V].double(), -1)
  mix = ref_sampler.ar_cfg_mix(lp_c, lp_u, gamma)
  ref = ref_sampler.ar_select(mix, G.double())
  xs, _ = helpers.sampler_kernel(
    logits_a=la.to(DEV), logits_b=lb.to(DEV), xt=torch.zeros(B, dtype=torch.long), mc_t=0.0,
    mc_s=0.0, vocab=V, seq_len=1, diffusion='ar', guidance='cfg', gamma=gamma, noise=G,
    select='gumbel')
  xs = xs.cpu()
  diff = xs != ref
  if bool(diff.any()):
    s = mix + G.double()
    gap = (s.gather(1, ref[:, None]) - s.gather(1, xs[:, None]))[diff[:, None]].abs()
    assert gap.max().item() <= 2e-5 * 20 and int(diff.sum()) <= 2, (int(diff.sum()), gap.max())


def test_ar_cfg_sample_follows_oracle_mix_of_own_logits():
  """`sample()` with parameterization=ar + guidance=cfg (prefix re-encode path):
  every id equals the oracle's AR-CFG selection applied to the CUDA path's own
  cond / uncond log-probs of that prefix (id parity downstream of identical
  logits), except isolated near-ties."""
  import dataclasses
  case = dataclasses.replace(cases.CASE_BY_NAME['ar_null'], name='ar_cfg', guidance='cfg',
                             gamma=2.0, length=32, batch=3)
  cfg, tok, sd, _ = helpers.case_state(case)
  model = ddg_b200.Diffusion(cfg, tok, device=DEV).load_state_dict(sd)
  model.round_bf16 = False
  G = cases.ar_gumbel(case)
  model.noise_fn = lambda i: G[:, i]
  model.ar_kv_cache = False
  x = model.sample()
  torch.cuda.synchronize()
  B, L = case.batch, case.length
  cond = torch.full((B,), case.condition, dtype=torch.long, device=DEV)
  lc = model.backbone.forward(x, None, cond).float().cpu()
  lu = model.backbone.forward(x, None, torch.full_like(cond, case.num_classes)).float().cpu()
  ok = 0
  for i in range(L - 1):
    mix = ref_sampler.ar_cfg_mix(torch.log_softmax(lc[:, i].double(), -1),
                                 torch.log_softmax(lu[:, i].double(), -1), case.gamma)
    ok += int((ref_sampler.ar_select(mix, G[:, i].double()) == x[:, i + 1].cpu()).sum())
  assert ok >= B * (L - 1) - 2, (ok, B * (L - 1))
