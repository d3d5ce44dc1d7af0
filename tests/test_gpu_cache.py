"""Per-sequence logits cache of the MDLM loop (the per-sequence form of the
reference's `sampling.use_cache`, diffusion.py:1371-1374) and the dynamic-batch
("live sequences") variants of the kernels it is built from."""
import dataclasses

import pytest
import torch

import ddg_b200
import helpers
from ddg_b200 import _lib
from golden import cases

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_seq_compact_orders_and_clears_flags():
  g = torch.Generator().manual_seed(3)
  for n in (1, 37, 512, 1024, 2500):
    flags = (torch.rand(n, generator=g) < 0.3).to(torch.int32).to(DEV)
    want = torch.nonzero(flags).flatten().to(torch.int32)
    seq_map = torch.full((n,), -1, dtype=torch.int32, device=DEV)
    count = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check(_lib.lib.ddg_seq_compact(flags.data_ptr(), n, seq_map.data_ptr(), count.data_ptr(),
                                        0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert int(count) == want.numel()
    assert torch.equal(seq_map[:want.numel()], want)
    assert int(flags.sum()) == 0


def _model(case):
  cfg, tok, sd, _ = helpers.case_state(case)
  return ddg_b200.Diffusion(cfg, tok, device=DEV).load_state_dict(sd), tok


@pytest.mark.parametrize('hidden,n_heads,L', [(128, 2, 32), (768, 12, 128)])
def test_dynamic_batch_forward_equals_full_forward_bitwise(hidden, n_heads, L):
  """A forward over the sequences `seq_map[0..n_live)` only (compact activations,
  every launch sized for the whole batch, live count read on the device) writes, for
  those sequences, exactly the logits and log-softmax partials of the full forward,
  and leaves every other row of the buffers untouched."""
  case = dataclasses.replace(cases.CASE_BY_NAME['mdlm_null'], name='dyn', hidden=hidden,
                             n_heads=n_heads, length=L, batch=40 if hidden == 768 else 12,
                             n_blocks=2, cond_dim=64)
  model, tok = _model(case)
  bb = model.backbone
  B, V = case.batch, tok.vocab_size
  g = torch.Generator().manual_seed(5)
  ids = torch.randint(0, V, (B, L), generator=g).to(DEV)
  mod = ddg_b200.ModSpec(n_mod=1, seqs_per_mod=B, sigma=torch.zeros(1, device=DEV))
  full = bb.alloc_logits(B * L)
  full_st = bb.alloc_stats(B * L)
  bb.run(ids, B, L, mod, full, stats_out=full_st, stats_mask_col=tok.mask_token_id)
  torch.cuda.synchronize()
  for picks in ([0], [3, 7, 8], list(range(B)), [B - 1], []):
    out = torch.full_like(full, 7.0)
    out_st = torch.full_like(full_st, -3.0)
    seq_map = torch.zeros(B, dtype=torch.int32, device=DEV)
    seq_map[:len(picks)] = torch.tensor(picks, dtype=torch.int32)
    n_live = torch.tensor([len(picks)], dtype=torch.int32, device=DEV)
    bb.run(ids, B, L, mod, out, stats_out=out_st, stats_mask_col=tok.mask_token_id,
           live=(seq_map, n_live))
    torch.cuda.synchronize()
    o3, f3 = out.view(B, L, -1), full.view(B, L, -1)
    s3, fs3 = out_st.view(-1, B, L, 2), full_st.view(-1, B, L, 2)
    for b in range(B):
      if b in picks:
        assert torch.equal(o3[b, :, :V], f3[b, :, :V]), (picks, b)
        assert torch.equal(s3[:, b], fs3[:, b]), (picks, b)
      else:
        assert bool((o3[b] == 7.0).all()) and bool((s3[:, b] == -3.0).all()), (picks, b)


@pytest.mark.parametrize('name,L,steps', [('mdlm_null', 32, 24), ('mdlm_null', 128, 200),
                                          ('bert_mdlm_null', 32, 40)])
def test_cached_loop_ids_equal_full_recompute(name, L, steps):
  """`sample()` with the per-sequence logits cache == `sample()` recomputing every
  sequence at every step, bit for bit, under the same Philox key; and with many
  steps the cache skips most of the work (counted from the flags)."""
  case = dataclasses.replace(cases.CASE_BY_NAME[name], name='cache', length=L, batch=24,
                             steps=steps)
  outs = {}
  for cache in (True, False):
    model, tok = _model(case)
    model.logits_cache = cache
    model.config.sampling.steps = steps
    a = model.sample()
    b = model.sample()             # second call: graph replay with a fresh prior
    torch.cuda.synchronize()
    outs[cache] = (a.cpu(), b.cpu())
    assert model._loop.cached == ('seq' if cache else None)
    assert int((a == tok.mask_token_id).sum()) == 0    # MDLM ends fully unmasked (eps = 1e-5)
  assert torch.equal(outs[True][0], outs[False][0])
  assert torch.equal(outs[True][1], outs[False][1])
  assert not torch.equal(outs[True][0], outs[True][1])


def test_cache_counts_live_sequences():
  """Drives the captured loop step by step and reads the device-side live count: the
  first step runs every sequence, later steps only those the sampler changed."""
  case = dataclasses.replace(cases.CASE_BY_NAME['mdlm_null'], name='cnt', length=32, batch=64,
                             steps=400)
  model, tok = _model(case)
  model.config.sampling.steps = 400
  B, L = 64, 32
  xt = model._sample_prior(B, L)
  sig, mct, mcs = model._step_scalars(400, 1e-5)
  loop = model._get_loop(B, L, None, None)
  assert loop.cached == 'seq'
  loop.prepare(xt, sig, mct, mcs, None)
  lives, prev = [], loop.xt.clone()
  for i in range(400):
    loop.replay(1)
    torch.cuda.synchronize()
    lives.append(int(loop.n_live))
    cur = loop.xt.clone()
    changed = (cur != prev).any(dim=1)
    if i + 1 < 400:   # what the NEXT step must recompute == sequences changed by this one
      assert int(loop.changed.sum()) == int(changed.sum())
    prev = cur
  assert lives[0] == B
  assert sum(lives) < 0.25 * 400 * B, sum(lives)     # L / steps = 0.08 changes per sequence-step
  assert int((loop.xt == tok.mask_token_id).sum()) == 0


def test_time_conditioned_and_guided_loops_do_not_cache_per_sequence():
  for name in ('udlm_null', 'mdlm_cfg_g2', 'd3pm_abs_submask'):
    case = cases.CASE_BY_NAME[name]
    model, tok = _model(case)
    model.sample()
    assert model._loop is None or model._loop.cached is None, name


@pytest.mark.parametrize('name', ['udlm_use_cache', 'udlm_cfg_use_cache'])
def test_use_cache_with_time_conditioning_matches_reference_rule(name):
  """`sampling.use_cache` + time conditioning: the reference keeps the
  log-probs of the previous step -- computed at an OLDER sigma -- while no token of the
  batch changes (diffusion.py:1371-1374; rule pinned to the reference's own trace in
  tests/test_oracle_pins.py).  The captured loop does it on the device (the forward of a
  step runs on all sequences if the previous step changed a token, on none otherwise);
  the eager loop does it on the host like the reference.  Same Philox key -> the two
  must agree bit for bit, both must skip forwards, and switching the cache off must
  change the result (the stale-sigma log-probs matter)."""
  case = cases.CASE_BY_NAME[name]
  steps = 200                                    # small moves per step: many unchanged steps
  outs, n_fwd = {}, {}
  for mode in ('graph', 'eager'):
    model, tok = _model(case)
    model.config.sampling.steps = steps
    model.use_cuda_graph = mode != 'eager'
    calls = {'n': 0}
    orig = model.backbone.run

    def counted(*a, _orig=orig, **k):
      calls['n'] += 1
      return _orig(*a, **k)
    if mode == 'eager':
      model.backbone.run = counted
    outs[mode] = model.sample().cpu()
    n_fwd[mode] = calls['n']
  assert torch.equal(outs['graph'], outs['eager'])
  assert 0 < n_fwd['eager'] < 0.8 * steps, n_fwd        # the host rule skipped forwards
  # the captured loop skips exactly the same steps: drive it step by step and read the
  # device-side live count (n_fwd * B sequences, or 0)
  model, tok = _model(case)
  model.config.sampling.steps = steps
  B, L = case.batch, case.length
  model._call_count = 1                                  # the key sample() used above
  xt = model._sample_prior(B, L)
  sig, mct, mcs = model._step_scalars(steps, 1e-5)
  guidance = model.config.get('guidance')
  cond = (torch.full((B,), guidance.condition, dtype=torch.long, device=DEV)
          if guidance is not None else None)
  loop = model._get_loop(B, L, None if guidance is None else guidance.method, guidance)
  assert loop.cached == 'batch'
  loop.prepare(xt, sig, mct, mcs, cond)
  lives = []
  for _ in range(steps):
    loop.replay(1)
    torch.cuda.synchronize()
    lives.append(int(loop.n_live))
  assert set(lives) <= {0, loop.n_fwd * B}
  assert sum(1 for v in lives if v) == n_fwd['eager'], (lives, n_fwd)
  assert torch.equal(loop.xt.cpu(), outs['graph'])
