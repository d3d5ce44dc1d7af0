"""GPU parity of the residual GEMM with the following LayerNorm + modulate fused into
its epilogue (`ddg_gemm` with `ln_weight`, reference models/dit.py:312-321, 256-259)
against the two separate launches it replaces (`ddg_gemm` GATE_RESIDUAL, then
`ddg_ln_modulate`) and against torch fp32.

The residual rows must be bit-identical to the unfused GEMM (same arithmetic).  The
normalised rows use the same operation order as `ddg_ln_modulate`, but mean and variance
are combined from per-128-column partials, so rstd can differ in its last ulp: the 16-bit
outputs may differ by one unit in the last place on a small fraction of the entries.
"""
import pytest
import torch

import ddg_b200
from ddg_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _work(M, N):
  return torch.zeros(int(_lib.lib.ddg_gemm_ln_work_bytes(M, N)), dtype=torch.uint8, device=DEV)


def _residual_gemm(a, w, x, bias, gate, gate_stride, rpm, *, ln=None, work=None, M=None,
                   live=None, seq_len=1, residual=None):
  """x += round(gate * round(a @ w^T + bias)); ln = (weight, shift, scale, stride, dtype, out)."""
  Mfull, K = a.shape
  N = w.shape[0]
  g = _lib.GemmArgs(
    a=a.data_ptr(), w=w.data_ptr(), out=x.data_ptr(), bias=_lib.ptr(bias), gate=_lib.ptr(gate),
    rope_cos=0, rope_sin=0, lda=a.stride(0), ldw=w.stride(0), ld_out=x.stride(0),
    gate_stride=gate_stride, rows_per_mod=rpm, M=M or Mfull, N=N, K=K,
    dtype=_lib.DTYPE_CODE[a.dtype], epilogue=_lib.EPI_GATE_RESIDUAL, seq_len=seq_len,
    head_dim=64, rope_cols=0, live_seqs=_lib.ptr(live), residual=_lib.ptr(residual))
  if ln is not None:
    lw, sh, sc, stride, dt, out = ln
    g.ln_weight, g.ln_shift, g.ln_scale = lw.data_ptr(), _lib.ptr(sh), _lib.ptr(sc)
    g.ln_mod_stride, g.ln_dtype, g.ln_out, g.ln_work = stride, _lib.DTYPE_CODE[dt], out.data_ptr(), work.data_ptr()
  _lib.check(_lib.lib.ddg_gemm(g, _lib.stream_ptr()))


def _ln(x, rows, lw, sh, sc, stride, rpm, dt, out):
  N = x.shape[1]
  _lib.check(_lib.lib.ddg_ln_modulate(
    x.data_ptr(), rows, N, 0, rows, 0, lw.data_ptr(), _lib.ptr(sh), _lib.ptr(sc), stride, rpm,
    _lib.DTYPE_CODE[dt], out.data_ptr(), _lib.stream_ptr()))


def _ulp_report(a, b, dt):
  """fraction of entries that differ, and the largest difference in units of the last place of
  the larger magnitude.  Differences below 1e-5 absolute are not counted in the second figure:
  where `y * (1 + scale) + shift` cancels to almost nothing, the last-ulp freedom of rstd is many
  ulps of the tiny result."""
  fa, fb = a.float(), b.float()
  d = (fa - fb).abs()
  eps = 2.0 ** -10 if dt == torch.float16 else 2.0 ** -7
  ulp = torch.maximum(fa.abs(), fb.abs()) * eps
  worst = torch.where(d > 1e-5, d / ulp.clamp_min(1e-30), torch.zeros_like(d)).max()
  return float((d != 0).float().mean()), float(worst)


CASES = [
  # M, N, K, rows_per_mod, block dtype, LayerNorm dtype, gate, modulate, bias
  (96, 128, 64, 32, torch.float16, torch.float16, True, True, True),
  (300, 256, 128, 1 << 40, torch.bfloat16, torch.bfloat16, False, False, False),
  (1000, 384, 256, 32, torch.float16, torch.float16, True, True, False),
  (4096 + 77, 768, 768, 128, torch.float16, torch.float16, True, True, False),   # attn-out
  (9000, 768, 3072, 128, torch.float16, torch.bfloat16, True, True, True),       # last fc2 -> final norm
  (2048, 1024, 512, 64, torch.bfloat16, torch.bfloat16, True, True, True),
  (40000, 768, 768, 1 << 40, torch.float16, torch.float16, True, True, True),    # > one wave of tiles
  (20000, 512, 512, 1280, torch.bfloat16, torch.bfloat16, True, True, False),    # two N tiles per row block
  (163840, 512, 2048, 40960, torch.bfloat16, torch.bfloat16, True, True, True),  # exact-CBG classifier chunk
]


@pytest.mark.parametrize('M,N,K,rpm,dt,ln_dt,with_gate,with_mod,with_bias', CASES)
def test_fused_layernorm_epilogue_matches_separate_launches(M, N, K, rpm, dt, ln_dt, with_gate,
                                                            with_mod, with_bias):
  g = torch.Generator().manual_seed(M + N + K)
  n_mod = 1 if rpm > M else (M + rpm - 1) // rpm
  a = torch.randn(M, K, generator=g).to(dt).to(DEV)
  w = (torch.randn(N, K, generator=g) * 0.05).to(dt).to(DEV)
  bias = None
  if with_bias:
    bias = torch.zeros((N + 255) // 256 * 256, dtype=dt, device=DEV)
    bias[:N] = (torch.randn(N, generator=g) * 0.2).to(dt).to(DEV)
  gate_tab = torch.randn(n_mod, 2 * N, generator=g).to(dt).to(DEV)
  gate = gate_tab[:, N:] if with_gate else None
  lw = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
  mod = torch.randn(n_mod, 3 * N, generator=g).to(ln_dt).to(DEV)
  sh = mod[:, :N] if with_mod else None
  sc = mod[:, N:2 * N] if with_mod else None
  # rows with a large common offset as well: the shifted partial sums must not cancel
  x0 = (torch.randn(M, N, generator=g) * 1.5 + torch.randn(M, 1, generator=g) * 4).to(DEV)
  work = _work(M, N)

  x_ref = x0.clone()
  _residual_gemm(a, w, x_ref, bias, gate, gate_tab.stride(0), rpm)
  ref = torch.empty(M, N, dtype=ln_dt, device=DEV)
  _ln(x_ref, M, lw, sh, sc, mod.stride(0), rpm, ln_dt, ref)

  for rep in range(3):     # later launches find the counters / parities the earlier ones left behind
    x = x0.clone()
    out = torch.full((M, N), float('nan'), dtype=ln_dt, device=DEV)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _residual_gemm(a, w, x, bias, gate, gate_tab.stride(0), rpm,
                   ln=(lw, sh, sc, mod.stride(0), ln_dt, out), work=work)
    e1.record()
    torch.cuda.synchronize()
    # a wait that only ends by its safety time-out (0.15 s) means a broken exchange
    assert e0.elapsed_time(e1) < 60.0, f'launch took {e0.elapsed_time(e1):.1f} ms'
    assert torch.equal(x, x_ref), 'residual rows differ from the unfused GEMM'
    # scratch layout: per 128-row block a 64-byte counter slot + 128 x N/128 words; the counter
    # word is (epoch parity << 8) | departures: every owner has left again
    stride = 64 + 128 * (N // 128) * 8
    words = work.view(-1, stride)[:, :4].contiguous().view(torch.int32)
    assert int((words & 255).max()) == 0, 'counters not reset'
    assert not torch.isnan(out.float()).any()
    frac, worst = _ulp_report(out, ref, ln_dt)
    assert worst <= 1.01 and frac < 5e-3, (frac, worst)
  # and against torch fp32
  y = torch.nn.functional.layer_norm(x_ref, [N]) * lw
  if with_mod:
    idx = torch.arange(M, device=DEV) // min(rpm, M)
    y = y * sc.float()[idx] + sh.float()[idx]
  rtol, atol = (4e-3, 2e-3) if ln_dt == torch.float16 else (2e-2, 1.5e-2)
  torch.testing.assert_close(out.float(), y.to(ln_dt).float(), rtol=rtol, atol=atol)


def test_fused_layernorm_launches_over_different_row_counts_share_the_scratch():
  """Exact D-CBG runs the classifier over chunks of different sizes with one workspace: the
  scratch layout must not depend on M (a launch over fewer rows then one over all of them)."""
  M, N, K = 20000, 768, 768
  g = torch.Generator().manual_seed(11)
  a = torch.randn(M, K, generator=g).half().to(DEV)
  w = (torch.randn(N, K, generator=g) * 0.05).half().to(DEV)
  lw = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
  x0 = torch.randn(M, N, generator=g).to(DEV)
  work = _work(M, N)
  x_ref = x0.clone()
  _residual_gemm(a, w, x_ref, None, None, 0, 1 << 40)
  ref = torch.empty(M, N, dtype=torch.float16, device=DEV)
  _ln(x_ref, M, lw, None, None, 0, 1 << 40, torch.float16, ref)
  for rows in (M, 5120, M, 256, 7777, M):
    x = x0.clone()
    out = torch.zeros(M, N, dtype=torch.float16, device=DEV)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _residual_gemm(a, w, x, None, None, 0, 1 << 40, ln=(lw, None, None, 0, torch.float16, out),
                   work=work, M=rows)
    e1.record()
    torch.cuda.synchronize()
    assert e0.elapsed_time(e1) < 60.0, f'{rows} rows: launch took {e0.elapsed_time(e1):.1f} ms'
    assert torch.equal(x[:rows], x_ref[:rows]) and torch.equal(x[rows:], x0[rows:])
    frac, worst = _ulp_report(out[:rows], ref[:rows], torch.float16)
    assert worst <= 1.01 and frac < 5e-3, (rows, frac, worst)


@pytest.mark.parametrize('fused', [False, True])
def test_residual_input_buffer_is_left_untouched(fused):
  """out = residual + delta with a separate residual input (the first-order D-CBG forward keeps
  every block's input rows this way): same values as the in-place form, input unchanged."""
  M, N, K = 3000, 768, 768
  g = torch.Generator().manual_seed(21)
  a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(DEV)
  w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(DEV)
  gate = torch.randn(1, N, generator=g).to(torch.bfloat16).to(DEV)
  lw = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
  x0 = torch.randn(M, N, generator=g).to(DEV)
  work = _work(M, N)

  def ln_args(out):
    return (lw, None, None, 0, torch.bfloat16, out) if fused else None
  x_inplace = x0.clone()
  y_inplace = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
  _residual_gemm(a, w, x_inplace, None, gate, N, 1 << 40, ln=ln_args(y_inplace), work=work)
  res = x0.clone()
  x_out = torch.full((M, N), float('nan'), device=DEV)
  y_out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
  _residual_gemm(a, w, x_out, None, gate, N, 1 << 40, ln=ln_args(y_out), work=work, residual=res)
  torch.cuda.synchronize()
  assert torch.equal(res, x0)
  assert torch.equal(x_out, x_inplace)
  assert torch.equal(y_out.view(torch.int16), y_inplace.view(torch.int16))


def test_fused_layernorm_over_dynamic_batch():
  """live_seqs: only the first n_live sequences exist; rows beyond them stay untouched."""
  L, n_seq, N, K = 32, 40, 768, 768
  M = L * n_seq
  g = torch.Generator().manual_seed(5)
  a = torch.randn(M, K, generator=g).half().to(DEV)
  w = (torch.randn(N, K, generator=g) * 0.05).half().to(DEV)
  gate_tab = torch.randn(1, N, generator=g).half().to(DEV)
  lw = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
  mod = torch.randn(1, 2 * N, generator=g).half().to(DEV)
  x0 = torch.randn(M, N, generator=g).to(DEV)
  work = _work(M, N)
  for n_live in (0, 7, 40):
    live = torch.tensor([n_live], dtype=torch.int32, device=DEV)
    x = x0.clone()
    out = torch.zeros(M, N, dtype=torch.float16, device=DEV)
    _residual_gemm(a, w, x, None, gate_tab, N, 1 << 40,
                   ln=(lw, mod[:, :N], mod[:, N:], 2 * N, torch.float16, out), work=work,
                   live=live, seq_len=L)
    torch.cuda.synchronize()
    rows = n_live * L
    x_ref = x0.clone()
    ref = torch.zeros(M, N, dtype=torch.float16, device=DEV)
    if rows:
      _residual_gemm(a[:rows], w, x_ref, None, gate_tab, N, 1 << 40, M=rows)
      _ln(x_ref, rows, lw, mod[:, :N], mod[:, N:], 2 * N, 1 << 40, torch.float16, ref)
    torch.cuda.synchronize()
    assert torch.equal(x, x_ref)
    assert torch.equal(out[rows:], ref[rows:])
    frac, worst = _ulp_report(out[:max(rows, 1)], ref[:max(rows, 1)], torch.float16)
    assert worst <= 1.01 and frac < 5e-3, (n_live, frac, worst)


def test_fused_layernorm_rejects_unsupported_shapes():
  a = torch.zeros(64, 64, dtype=torch.float16, device=DEV)
  w = torch.zeros(192, 64, dtype=torch.float16, device=DEV)    # N % 128 != 0
  x = torch.zeros(64, 192, device=DEV)
  lw = torch.ones(192, device=DEV)
  out = torch.zeros(64, 192, dtype=torch.float16, device=DEV)
  with pytest.raises(RuntimeError):
    _residual_gemm(a, w, x, None, None, 0, 1 << 40,
                   ln=(lw, None, None, 0, torch.float16, out), work=_work(64, 192))


@pytest.mark.parametrize('hidden,n_heads,L,B', [(128, 2, 32, 6), (768, 12, 128, 40)])
def test_forward_with_fused_layernorm_matches_unfused_forward(hidden, n_heads, L, B):
  """Whole denoiser forward, fused vs separate LayerNorm launches: logits agree to bf16
  rounding noise (the 16-bit LayerNorm outputs differ by at most one ulp on a few entries)."""
  V = 1000
  cfg = ddg_b200.compose([
    'parameterization=d3pm', 'diffusion=uniform', 'time_conditioning=True', 'T=0', 'model=small',
    f'model.hidden_size={hidden}', f'model.n_heads={n_heads}', 'model.n_blocks=3',
    f'model.length={L}', 'model.cond_dim=128', 'guidance=cfg', 'guidance.gamma=2',
    '+data.num_classes=2'])
  sd = ddg_b200.weights.random_state_dict(cfg, V, seed=3)
  net = ddg_b200.DIT(cfg, vocab_size=V, device=DEV).load_state_dict(sd)
  ids = torch.randint(0, V, (B, L), device=DEV)
  sigma = torch.linspace(0.2, 3.0, B, device=DEV)
  cond = torch.randint(0, 3, (B,), device=DEV)
  net.fuse_ln = True
  fused = net.forward(ids, sigma, cond).float().clone()
  net.fuse_ln = False
  plain = net.forward(ids, sigma, cond).float().clone()
  torch.cuda.synchronize()
  err = (fused - plain).abs()
  scale = plain.abs().mean()
  assert float(err.mean() / scale) < 2e-3, float(err.mean() / scale)
  assert float(err.max() / scale) < 0.2, float(err.max() / scale)
