"""CPU tests of the host-side logic: config composition (the reference's
hydra override surface), schedules, weight naming, sharding, Philox helper."""
import numpy as np
import pytest
import torch

import ddg_b200
import helpers
from ddg_b200 import dist as ddist
from golden import cases
from oracle import ref_sampler


def test_compose_overrides_like_hydra():
  cfg = ddg_b200.compose(['parameterization=d3pm', 'diffusion=uniform',
                          'guidance=cfg', 'guidance.gamma=2', 'model=small',
                          'model.length=128', '+data.num_classes=2',
                          'sampling.steps=1000', 'time_conditioning=True',
                          'sampling.batch_size=512'])
  assert cfg.guidance.method == 'cfg' and cfg.guidance.gamma == 2
  assert cfg.model.hidden_size == 768 and cfg.model.n_blocks == 12
  assert cfg.model.length == 128 and cfg.data.num_classes == 2
  assert cfg.sampling.steps == 1000 and cfg.time_conditioning is True
  cfg = ddg_b200.compose(['guidance=cbg', 'classifier_model=tiny-classifier',
                          'data=qm9', 'guidance.use_approx=True'])
  assert cfg.classifier_model.hidden_size == 512
  assert cfg.classifier_model.num_classes == 2 and cfg.guidance.use_approx
  assert ddg_b200.compose([]).guidance is None
  assert ddg_b200.compose(['training.guidance=null']).training.guidance is None
  with pytest.raises(KeyError):
    ddg_b200.compose(['model=huge'])


def test_validate_configuration():
  bad = ddg_b200.compose(['parameterization=subs', 'diffusion=uniform'])
  with pytest.raises(AssertionError):
    ddg_b200.config.validate(bad)
  ddg_b200.config.validate(ddg_b200.compose(['parameterization=d3pm', 'diffusion=uniform', 'T=1000']))


@pytest.mark.parametrize('steps,T', [(4, 0), (128, 0), (1000, 0), (7, 1000)])
def test_step_scalars_match_oracle(steps, T):
  from ddg_b200.diffusion import Diffusion
  fake = Diffusion.__new__(Diffusion)
  fake.T = T
  fake.noise = ddg_b200.noise_schedule.LogLinearNoise()
  sig, mct, mcs = Diffusion._step_scalars(fake, steps, 1e-5)
  s2, m2, m3 = ref_sampler.step_scalars(steps, 1e-5, T)
  torch.testing.assert_close(sig, s2, rtol=0, atol=0)
  torch.testing.assert_close(mct, m2, rtol=0, atol=0)
  torch.testing.assert_close(mcs, m3, rtol=0, atol=0)
  assert bool((mcs < mct).all())


def test_random_state_dict_names_and_shapes():
  cfg = ddg_b200.compose(['parameterization=d3pm', 'diffusion=uniform', 'guidance=cfg',
                          '+data.num_classes=2', 'model=tiny'])
  sd = ddg_b200.weights.random_state_dict(cfg, 35, seed=1)
  assert sd['backbone.vocab_embed.embedding'].shape == (35, 512)
  assert sd['backbone.blocks.7.adaLN_modulation.weight'].shape == (3072, 128)
  assert sd['backbone.cond_map.embedding_table.weight'].shape == (3, 128)
  assert sd['backbone.output_layer.linear.weight'].shape == (35, 512)
  sd2 = ddg_b200.weights.random_state_dict(cfg, 35, seed=1)
  assert all(torch.equal(sd[k], sd2[k]) for k in sd)


def test_philox_helper_known_answer():
  """Philox4x32-10 known-answer vectors (Random123 kat_vectors): counter 0 /
  key 0 and counter, key = all ones."""
  def words(ctr, key):
    # one call: rows=1 row index provides c1,c2; emulate by direct loop below
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
      p0, p1 = M0 * c[0], M1 * c[2]
      c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
      k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c
  assert words([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
  assert words([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
  # the vectorised helper agrees with the scalar loop
  seed, step, rows, V = 0x0123456789ABCDEF, 5, 3, 10
  U = helpers.philox_uniform(seed, step, rows, V)
  for r in range(rows):
    for j4 in range(3):
      w = words([j4, r, 0, step], [seed & 0xFFFFFFFF, seed >> 32])
      for e in range(4):
        j = 4 * j4 + e
        if j < V:
          assert U[r, j] == np.float32((w[e] >> 8) * 2.0 ** -24)


def test_shard_bounds_cover_batch():
  for total in (512, 513, 7):
    for world in (1, 2, 4, 8):
      spans = [ddist.shard_bounds(total, world, r) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
  assert len({ddist.rank_seed(1, r) for r in range(8)}) == 8


def test_oracle_mdlm_resolves_all_masks():
  """Domain property on the CPU oracle: the log-linear MDLM sampler leaves no
  mask after the last step (move_chance_s ~ 1e-5)."""
  case = cases.CASE_BY_NAME['mdlm_null']
  oracle = helpers.make_oracle(case)
  xt = cases.prior_ids(case)
  g = torch.Generator().manual_seed(0)
  V = cases.vocab_size(case)
  out = oracle.sample(xt, lambda i: torch.rand(case.batch, case.length, V, generator=g), 16)
  assert (out == cases.Tok(case.vocab).mask_token_id).float().mean().item() < 0.05


def test_wide_group_columns_partition_the_vocabulary():
  """race_group = 64: the strided 64-column groups cover every column exactly once."""
  import numpy as np
  from oracle import ref_sampler
  for V in (30522, 4096, 2051, 40):
    cols = ref_sampler.wide_group_columns(V)
    live = cols[cols >= 0]
    assert live.size == V and np.array_equal(np.sort(live), np.arange(V))
    assert cols.shape[1] == 64
  c = ref_sampler.wide_group_columns(30522)
  assert c[0, :9].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 1024]      # (s, e) order, 128 vectors apart
  assert c[129, 0] == 8 * (1 + 128 * 8)                            # t = 1, sweep 1


@pytest.mark.parametrize('group', [8, 64])
def test_group_race_restatement_is_a_categorical_draw(group):
  """oracle.group_race_select (the CPU restatement of the production
  sampler's hierarchical exponential race) reproduces the categorical
  distribution it is given, and is a pure function of (seed, step, row)."""
  import torch
  from oracle import ref_sampler
  V, n = (83, 30000) if group == 8 else (2051, 30000)   # ragged last group / vector
  g = torch.Generator().manual_seed(0)
  q = torch.softmax(torch.randn(V, generator=g) * 2, 0)
  q[5] = 0.0
  r1 = ref_sampler.group_race_select(q[None].repeat(n, 1), 12345, 7, group=group)
  r2 = ref_sampler.group_race_select(q[None].repeat(n, 1), 12345, 7, group=group)
  r3 = ref_sampler.group_race_select(q[None].repeat(n, 1), 12345, 8, group=group)
  assert torch.equal(r1['ids'], r2['ids']) and not torch.equal(r1['ids'], r3['ids'])
  assert int(r1['ids'].max()) < V
  c = torch.bincount(r1['ids'], minlength=V).double()
  assert c[5] == 0
  qn = (q / q.sum()).double()
  z = (c - qn * n) / (qn * n * (1 - qn)).clamp_min(1).sqrt()
  assert z.abs().max().item() < 4.5, z


def test_unsupported_model_length_is_reported_at_construction():
  """The reference's default `model.length=1024` is outside the attention kernels' range
  (<= 256 tokens, multiples of 16): the constructor says so instead of the first forward."""
  import ddg_b200
  from ddg_b200 import dit
  cfg = ddg_b200.compose(['model=small', '+data.num_classes=2'])
  assert cfg.model.length == 1024
  with pytest.raises(NotImplementedError, match='model.length'):
    dit.DIT(cfg, 100)
  cfg = ddg_b200.compose(['model=small', 'model.length=40', '+data.num_classes=2'])
  with pytest.raises(NotImplementedError, match='multiples of 16'):
    dit.DIT(cfg, 100)
  cfg = ddg_b200.compose(['model=small', 'model.length=128', '+data.num_classes=2'])
  assert dit.DIT(cfg, 100).hidden == 768
