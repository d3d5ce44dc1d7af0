"""Where the joules of one reverse step go.  Each kernel family of the bench step is run
back to back for ~2 s (after a pre-heat) while NVML board power and the SM clock are
sampled: time per launch, average power, energy per launch and per step.  Under the
board's power cap the step time is (sum of energies) / cap, not a sum of roofline times.
`python scripts/power_probe.py`"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pynvml
import torch

import helpers
from ddg_b200 import _lib

DEV = 'cuda'
SECONDS = float(os.environ.get('SUSTAIN_S', '2.0'))


class Sampler(threading.Thread):
  def __init__(self, h):
    super().__init__(daemon=True)
    self.h, self.stop, self.p, self.c = h, False, [], []

  def run(self):
    while not self.stop:
      self.p.append(pynvml.nvmlDeviceGetPowerUsage(self.h) / 1e3)
      self.c.append(pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM))
      time.sleep(0.02)


def sustained(fn, h):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    fn()
  e1.record()
  torch.cuda.synchronize()
  per = e0.elapsed_time(e1) / 5
  n = max(10, int(SECONDS * 1e3 / per))
  for _ in range(n // 2):
    fn()
  torch.cuda.synchronize()
  smp = Sampler(h)
  e0.record()
  smp.start()
  for _ in range(n):
    fn()
  e1.record()
  torch.cuda.synchronize()
  smp.stop = True
  smp.join()
  ms = e0.elapsed_time(e1) / n
  k = len(smp.p) // 4                      # drop the ramp at the start of the window
  p = sorted(smp.p[k:]) or [0.0]
  c = sorted(smp.c[k:]) or [0]
  return ms, p[len(p) // 2], c[len(c) // 2]


def main():
  pynvml.nvmlInit()
  h = pynvml.nvmlDeviceGetHandleByIndex(0)
  cap = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
  idle = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
  print(f'enforced power limit {cap:.0f} W, idle {idle:.0f} W', flush=True)
  L, M, hid, H, V, ld = 128, 131072, 768, 12, 30522, 30528
  sp = _lib.stream_ptr()
  rows = []

  def gemm_case(name, N, K, epi, dt, launches):
    a = (torch.randn(M, K, device=DEV) * 0.7).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.03).to(dt)
    bias = (torch.randn((N + 255) // 256 * 256, device=DEV) * 0.1).to(dt)
    res = epi == _lib.EPI_GATE_RESIDUAL
    ldo = (N + 7) // 8 * 8
    out = torch.zeros(M, ldo, dtype=torch.float32 if res else dt, device=DEV)
    gate = (torch.randn(2, N, device=DEV) * 0.1).to(dt)
    cos = torch.randn(L, 32, device=DEV).to(dt)
    sin = torch.randn(L, 32, device=DEV).to(dt)
    stats = torch.empty((N + 127) // 128, M, 2, dtype=torch.float32, device=DEV) if N > 4096 else None
    g = _lib.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), bias=bias.data_ptr(),
                      gate=gate.data_ptr() if res else None, rope_cos=cos.data_ptr(), rope_sin=sin.data_ptr(),
                      lda=K, ldw=K, ld_out=ldo, gate_stride=N, rows_per_mod=M // 2, M=M, N=N, K=K,
                      dtype=_lib.DTYPE_CODE[dt], epilogue=epi, seq_len=L, head_dim=64,
                      rope_cols=2 * N // 3 if epi == _lib.EPI_QKV_ROPE else 0,
                      stats=stats.data_ptr() if stats is not None else None, stats_ld=M, stats_mask_col=-1)
    ms, p, c = sustained(lambda: _lib.check(_lib.lib.ddg_gemm(g, sp)), h)
    rows.append((name, launches, ms, p, c, 2.0 * M * N * K / ms / 1e9))

  gemm_case('gemm qkv+rope', 3 * hid, hid, _lib.EPI_QKV_ROPE, torch.float16, 12)
  gemm_case('gemm attn_out', hid, hid, _lib.EPI_GATE_RESIDUAL, torch.float16, 12)
  gemm_case('gemm fc1+gelu', 4 * hid, hid, _lib.EPI_BIAS_GELU, torch.float16, 12)
  gemm_case('gemm fc2', hid, 4 * hid, _lib.EPI_GATE_RESIDUAL, torch.float16, 12)
  gemm_case('gemm vocab', V, hid, _lib.EPI_STORE, torch.bfloat16, 1)
  torch.cuda.empty_cache()

  # LayerNorm + modulate
  x = torch.randn(M, hid, device=DEV)
  lnw = torch.ones(hid, device=DEV)
  mod = torch.randn(2, 2 * hid, device=DEV).half()
  o16 = torch.empty(M, hid, dtype=torch.float16, device=DEV)
  ms, p, c = sustained(lambda: _lib.check(_lib.lib.ddg_ln_modulate(
    x.data_ptr(), M, hid, 0, M, 0, lnw.data_ptr(), mod.data_ptr(), mod.data_ptr() + hid * 2, 2 * hid, M // 2,
    _lib.F16, o16.data_ptr(), sp)), h)
  rows.append(('ln_modulate', 25, ms, p, c, None))
  # attention
  qkv = torch.randn(M, 3 * hid, device=DEV).half()
  ms, p, c = sustained(lambda: _lib.check(_lib.lib.ddg_attention(
    qkv.data_ptr(), o16.data_ptr(), M // L, L, H, 64, 0, _lib.F16, sp)), h)
  rows.append(('attention', 12, ms, p, c, None))
  del qkv, x
  torch.cuda.empty_cache()
  # fused sampler (CFG, hierarchical race): 16384 rows per launch, scaled to the step's 65536
  R = 16384
  g = torch.Generator(device=DEV).manual_seed(0)
  la = torch.empty(R, ld, dtype=torch.bfloat16, device=DEV)
  lb = torch.empty(R, ld, dtype=torch.bfloat16, device=DEV)
  for t in (la, lb):
    for r0 in range(0, R, 4096):
      t[r0:r0 + 4096] = (torch.randn(4096, ld, device=DEV, generator=g) * 2.0).to(torch.bfloat16)
  st_a = torch.cat([helpers.make_stats(la[r0:r0 + 4096], V) for r0 in range(0, R, 4096)], 1).contiguous()
  st_b = torch.cat([helpers.make_stats(lb[r0:r0 + 4096], V) for r0 in range(0, R, 4096)], 1).contiguous()
  xt = torch.randint(0, V, (R,), device=DEV, generator=g)
  xs = torch.empty_like(xt)
  mct = torch.tensor([0.5], device=DEV)
  mcs = torch.tensor([0.499], device=DEV)
  a = _lib.SampleStepArgs(
    logits_a=la.data_ptr(), logits_b=lb.data_ptr(), clf=0, noise=0, xt=xt.data_ptr(),
    mc_t=mct.data_ptr(), mc_s=mcs.data_ptr(), step_ptr=0, seed_ptr=0, xs=xs.data_ptr(), q_out=0,
    rows=R, ld_a=ld, ld_b=ld, ld_clf=0, ld_noise=0, ld_q=0, seq_len=L, vocab=V,
    mc_seq_stride=0, mc_step_stride=0, logits_dtype=_lib.BF16, diffusion=_lib.UNIFORM,
    guidance=_lib.GUIDE_CFG, subs=0, mask_logit=0, mask_index=103, round_bf16=1,
    select=_lib.SELECT_RACE, fast_math=1, gamma=2.0, seed=1, step=3, stats_chunks=st_a.shape[0],
    stats_a=st_a.data_ptr(), stats_b=st_b.data_ptr(), stats_ld=R, race_group=64)
  ms, p, c = sustained(lambda: _lib.check(_lib.lib.ddg_sample_step(a, sp)), h)
  rows.append(('sampler (x4 rows)', 4, ms, p, c, None))

  tot_ms = tot_j = 0.0
  print(f'{"kernel":20s} {"n/step":>6s} {"ms":>8s} {"W":>6s} {"MHz":>6s} {"J/launch":>9s} {"ms/step":>8s} {"J/step":>7s}  TF/s')
  for name, n, ms, p, c, tf in rows:
    j = p * ms / 1e3
    tot_ms += n * ms
    tot_j += n * j
    print(f'{name:20s} {n:6d} {ms:8.3f} {p:6.0f} {c:6d} {j:9.3f} {n * ms:8.2f} {n * j:7.2f}  {"" if tf is None else round(tf)}')
  print(f'sum of sustained times {tot_ms:.2f} ms/step, energy {tot_j:.1f} J/step -> {tot_j / cap * 1e3:.2f} ms at the {cap:.0f} W cap')


if __name__ == '__main__':
  main()
