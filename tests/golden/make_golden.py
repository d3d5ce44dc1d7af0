"""Generates tests/golden/*.npz by running the UNMODIFIED reference sampler
(/root/reference, imported through oracle/refshim.py) on CPU with injected
prior and noise.  Run in the build container only:

    python tests/golden/make_golden.py [case ...]

The reference is on CPU, where every autocast in it is disabled, so these
vectors pin the fp32 math of the path (SURVEY.md 8a' item 7).

    python tests/golden/make_golden.py --amp [case ...]

writes `<case>_amp.npz` instead: the same unmodified reference code, but with
`torch.cuda.amp.autocast` mapped onto CPU autocast of the same dtype
(oracle/refshim.cuda_autocast_on_cpu), so every tensor is rounded where the
reference rounds it on a GPU (fp16 blocks, bf16 final layer, bf16 log-softmax
and exp).  These vectors pin the ROUNDING POINTS of the path; `*_dtype`
entries record the dtype each traced tensor had.

    python tests/golden/make_golden.py --gpu --out gpurun_out/golden_gpu [case ...]

(on the B200 box, from the staged copy baseline/_ref) writes `<case>_gpu.npz`:
the unmodified reference ON THE GPU under its real cuda autocast (cuBLAS fp16 /
bf16 GEMMs, flash-attn Triton rotary, SDPA) -- the pin SURVEY.md 8a' asks to
"confirm on the first gpurun".  The committed `*_gpu.npz` were produced that way.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from golden import cases  # noqa: E402
from oracle import ref_model, refshim  # noqa: E402

TOPK = 8


def build_reference_model(ref, case, device='cpu'):
  tok = cases.Tok(case.vocab)
  model, clf, cfg, _, _ = ref_model.build_reference(
    cases.config_dict(case), tok, case.seed, device=device,
    with_classifier=case.guidance == 'cbg')
  return model, clf, cfg


def summarise(name, t, out, big):
  t = t.detach().float().cpu()
  if not big:
    out[name] = t.numpy()
    return
  top = t.topk(TOPK, dim=-1)
  out[name + '_topv'] = top.values.numpy()
  out[name + '_topi'] = top.indices.numpy().astype(np.int32)
  out[name + '_sum'] = t.double().sum(-1).numpy()
  # a fixed strided slice of columns pins the bulk of the distribution too
  out[name + '_cols'] = t[..., ::997].numpy()


def run_case(ref, case, device='cpu'):
  torch.manual_seed(0)
  model, clf, cfg = build_reference_model(ref, case, device)
  big = case.vocab == 'bert'
  out = {}
  xt0 = cases.prior_ids(case).to(device)
  model._sample_prior = lambda *dims: xt0.clone()

  if case.parameterization == 'ar':
    G = cases.ar_gumbel(case).to(device)

    class FakeGumbel:
      def __init__(self, *a, **k):
        pass

      def sample(self, shape):
        assert tuple(shape) == tuple(G.shape), (shape, G.shape)
        return G.clone()
    real = torch.distributions.Gumbel
    torch.distributions.Gumbel = FakeGumbel
    try:
      with torch.no_grad():
        x = model.sample()
    finally:
      torch.distributions.Gumbel = real
    out['samples'] = x.cpu().numpy().astype(np.int32)
    # log-probs of the final sequence pin the causal forward
    with torch.no_grad():
      if case.guidance == 'cfg':
        cond = torch.full((case.batch,), case.condition, dtype=torch.long, device=device)
        summarise('log_probs_final_cond', model.forward(x, None, cond=cond), out, big)
        summarise('log_probs_final_uncond', model.forward(
          x, None, cond=torch.full_like(cond, case.num_classes)), out, big)
      else:
        summarise('log_probs_final', model.forward(x, None), out, big)
    return out

  if clf is not None:
    ref.classifier.Classifier.load_from_checkpoint = classmethod(
      lambda cls, *a, **k: clf)

  trace = []
  step_counter = {'i': 0}
  real_rand_like = torch.rand_like

  def fake_rand_like(t, *a, **k):
    U = cases.step_noise(case, step_counter['i'])
    assert U.shape == t.shape, (U.shape, t.shape)
    return U.to(t.device, t.dtype)

  def wrap(fn_name):
    orig = getattr(model, fn_name)

    def wrapped(*a, **k):
      xt = k['xt']
      xs, q, cache = orig(*a, **k)
      rec = {'xt': xt.clone(), 'xs': xs.clone(), 'q_xs': q.clone(),
             'cached': k.get('cache') is not None,
             'dtypes': {'q_xs': str(q.dtype), **{
               ck: str(cv.dtype) for ck, cv in cache.items() if cv is not None}},
             'mc_t': k['move_chance_t'].reshape(-1)[0].clone(),
             'mc_s': k['move_chance_s'].reshape(-1)[0].clone(),
             'sigma_t': k['time_conditioning'].reshape(-1)[0].clone()}
      for ck, cv in cache.items():
        if cv is not None:
          rec[ck] = cv.clone()
      trace.append(rec)
      step_counter['i'] += 1
      return xs, q, cache
    setattr(model, fn_name, wrapped)

  for fn in ('_ddpm_denoise', '_cfg_denoise', '_cbg_denoise'):
    wrap(fn)
  torch.rand_like = fake_rand_like
  try:
    with torch.no_grad():
      samples = model.sample()
  finally:
    torch.rand_like = real_rand_like
  assert len(trace) == case.steps, (len(trace), case.steps)
  out['samples'] = samples.cpu().numpy().astype(np.int32)
  out['xt0'] = xt0.cpu().numpy().astype(np.int32)
  for i, rec in enumerate(trace):
    out[f's{i}_xt'] = rec['xt'].cpu().numpy().astype(np.int32)
    out[f's{i}_xs'] = rec['xs'].cpu().numpy().astype(np.int32)
    if case.use_cache:   # 1 = the reference re-used the cached log-probs in this step
      out[f's{i}_cached'] = np.array(int(rec['cached']), dtype=np.int32)
    out[f's{i}_scalars'] = np.array(
      [rec['sigma_t'].item(), rec['mc_t'].item(), rec['mc_s'].item()],
      dtype=np.float32)
    summarise(f's{i}_q_xs', rec['q_xs'], out, big)
    for k in ('log_x_theta', 'log_x_theta_cond', 'log_x_theta_uncond',
              'classifier_log_prob'):
      if k in rec:
        summarise(f's{i}_{k}', rec[k], out, big)
  for k, v in trace[0]['dtypes'].items():
    out[f'{k}_dtype'] = np.array(v)
  return out


def main(argv):
  ref = refshim.load_reference()
  amp = '--amp' in argv
  gpu = '--gpu' in argv
  out_dir = HERE
  if '--out' in argv:
    out_dir = argv[argv.index('--out') + 1]
    argv = [a for i, a in enumerate(argv) if a != '--out' and argv[i - 1] != '--out']
    os.makedirs(out_dir, exist_ok=True)
  argv = [a for a in argv if a not in ('--amp', '--gpu')]
  names = argv or (cases.AMP_CASES if (amp or gpu) else [c.name for c in cases.CASES])
  for name in names:
    case = cases.CASE_BY_NAME[name]
    if gpu:      # the reference itself on the GPU under its REAL cuda autocast policy
      out = run_case(ref, case, device='cuda')
    elif amp:
      with refshim.cuda_autocast_on_cpu():
        out = run_case(ref, case)
    else:
      out = run_case(ref, case)
    suffix = '_gpu' if gpu else ('_amp' if amp else '')
    path = os.path.join(out_dir, f'{name}{suffix}.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB')


if __name__ == '__main__':
  main(sys.argv[1:])
