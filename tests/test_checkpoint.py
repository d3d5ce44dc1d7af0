"""Checkpoint ingestion (reference diffusion.py:194-213,
models/ema.py:93-101): Lightning `.ckpt` files written FROM THE REFERENCE MODULES
(state_dict + `ema.state_dict()` with perturbed shadow params, classifier ckpt)
must load into the CUDA path with the reference's semantics -- EMA weights unless
`eval.disable_ema`, raw weights otherwise."""
import copy

import pytest
import torch

import ddg_b200
import helpers
from golden import cases
from oracle import ref_dit, ref_model, refshim

needs_reference = pytest.mark.skipif(
  not refshim.reference_available(),
  reason='no copy of the reference reachable (baseline/_ref is staged by build())')

CASE = cases.CASE_BY_NAME['udlm_cbg']


@pytest.fixture(scope='module')
def ckpts(tmp_path_factory):
  out = tmp_path_factory.mktemp('ckpt')
  return ref_model.make_reference_checkpoints(
    cases.config_dict(CASE), cases.Tok(CASE.vocab), CASE.seed, str(out))


@needs_reference
def test_ema_shadow_params_map_onto_reference_parameter_order(ckpts):
  """`load_checkpoint_state` maps `ema['shadow_params'][i]` onto state_dict keys by
  order; the order must be the reference's `chain(backbone.parameters(),
  noise.parameters())` (diffusion.py:144-151).  Oracle: the reference's own
  `ema.copy_to` followed by `state_dict()`."""
  sd_ema = ddg_b200.weights.load_checkpoint_state(ckpts['diffusion'], use_ema=True)
  sd_raw = ddg_b200.weights.load_checkpoint_state(ckpts['diffusion'], use_ema=False)
  keys = [k for k in ckpts['ema_sd'] if k.startswith('backbone.')]
  assert set(sd_ema) == set(keys) == set(sd_raw)
  n_changed = 0
  for k in keys:
    assert torch.equal(sd_ema[k], ckpts['ema_sd'][k]), k
    assert torch.equal(sd_raw[k], ckpts['raw_sd'][k]), k
    n_changed += int(not torch.equal(sd_ema[k], sd_raw[k]))
  assert n_changed >= len(keys) - 1      # everything but the rotary buffer moved
  csd = ddg_b200.weights.load_checkpoint_state(ckpts['classifier'], use_ema=False,
                                               prefix='classifier_model.')
  assert set(csd) == set(ckpts['clf_sd'])
  for k, v in ckpts['clf_sd'].items():
    assert torch.equal(csd[k], v), k


@needs_reference
def test_shape_mismatch_in_shadow_params_is_reported(ckpts, tmp_path):
  ck = torch.load(ckpts['diffusion'], map_location='cpu', weights_only=False)
  ck['ema']['shadow_params'][3] = torch.zeros(3, 3)
  p = tmp_path / 'bad.ckpt'
  torch.save(ck, p)
  with pytest.raises(ValueError, match='EMA shadow param'):
    ddg_b200.weights.load_checkpoint_state(str(p), use_ema=True)


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize('disable_ema', [False, True])
def test_load_from_checkpoint_reproduces_reference_weights_on_gpu(ckpts, disable_ema):
  """`Diffusion.load_from_checkpoint` (main.py:39-47) / `Classifier.
  load_from_checkpoint`: logits of the CUDA forward == oracle forward on the EMA
  (resp. raw) weights, and not on the other set."""
  cfg, tok, _, _ = helpers.case_state(CASE)
  cfg = copy.deepcopy(cfg)
  cfg.eval.disable_ema = disable_ema
  model = ddg_b200.Diffusion.load_from_checkpoint(ckpts['diffusion'], tokenizer=tok,
                                                  config=cfg, device='cuda')
  B, L, V = 4, CASE.length, tok.vocab_size
  g = torch.Generator().manual_seed(3)
  ids = torch.randint(0, V, (B, L), generator=g)
  sigma = torch.rand(B, generator=g) * 3
  got = model._logits(ids.cuda(), sigma.cuda()).float().cpu()
  want_sd = ckpts['raw_sd'] if disable_ema else ckpts['ema_sd']
  other_sd = ckpts['ema_sd'] if disable_ema else ckpts['raw_sd']
  rnd = ref_dit.Rounding.gpu_denoiser()
  want = ref_dit.dit_forward(want_sd, ids, sigma, n_heads=CASE.n_heads, rounding=rnd)
  other = ref_dit.dit_forward(other_sd, ids, sigma, n_heads=CASE.n_heads, rounding=rnd)
  d = (got - want).abs()
  assert bool((d <= 0.06 + 0.03 * want.abs()).all()), d.max().item()
  assert d.mean().item() < 0.3 * (got - other).abs().mean().item()
  clf = ddg_b200.Classifier.load_from_checkpoint(ckpts['classifier'], tokenizer=tok,
                                                 config=cfg, device='cuda')
  lp = clf.get_log_probs(ids.cuda(), sigma.cuda()).cpu()
  ref = ref_dit.classifier_log_probs(ckpts['clf_sd'], ids, sigma, n_heads=CASE.n_heads,
                                     time_conditioning=True,
                                     rounding=ref_dit.Rounding.gpu_classifier())
  torch.testing.assert_close(lp, ref, rtol=3e-2, atol=3e-2)


@needs_reference
@pytest.mark.gpu
def test_sample_with_cbg_from_checkpoints_on_gpu(ckpts):
  """The documented quick-start path end to end: `load_from_checkpoint` +
  `guidance=cbg guidance.classifier_checkpoint_path=...` + `sample()`; the
  classifier is loaded once (reference reloads it per call, diffusion.py:1062-1068)."""
  cfg, tok, _, _ = helpers.case_state(CASE)
  cfg = copy.deepcopy(cfg)
  cfg.eval.disable_ema = False
  cfg.guidance.classifier_checkpoint_path = ckpts['classifier']
  model = ddg_b200.Diffusion.load_from_checkpoint(ckpts['diffusion'], tokenizer=tok,
                                                  config=cfg, device='cuda')
  out = model.sample()
  torch.cuda.synchronize()
  assert tuple(out.shape) == (CASE.batch, CASE.length) and out.dtype == torch.int64
  assert int(out.min()) >= 0 and int(out.max()) < tok.vocab_size
  first = model.classifier_model
  model.sample()
  assert model.classifier_model is first
