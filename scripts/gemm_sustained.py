"""Sustained (power-capped) throughput of the step's GEMM shapes, one shape x epilogue at a
time for ~2 s each after a pre-heat, with the SM clock sampled under load -- what a kernel
delivers INSIDE a long step rather than in a cold burst.  `python scripts/gemm_sustained.py`"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch

import ddg_b200
from ddg_b200 import _lib

DEV = 'cuda'
SECONDS = float(os.environ.get('SUSTAIN_S', '2.0'))


def sustained(fn, flops, h):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  # calibrate
  e0.record()
  for _ in range(5):
    fn()
  e1.record()
  torch.cuda.synchronize()
  per = e0.elapsed_time(e1) / 5
  n = max(10, int(SECONDS * 1e3 / per))
  # pre-heat half, measure half
  for _ in range(n // 2):
    fn()
  e0.record()
  clocks = []
  for i in range(n):
    fn()
    if i % max(1, n // 8) == 0:
      clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  clocks.sort()
  return ms, flops / ms / 1e9, clocks[len(clocks) // 2]


def main():
  pynvml.nvmlInit()
  h = pynvml.nvmlDeviceGetHandleByIndex(0)
  L = 128
  cases = [
    ('qkv store', 131072, 2304, 768, _lib.EPI_STORE, torch.float16),
    ('qkv rope', 131072, 2304, 768, _lib.EPI_QKV_ROPE, torch.float16),
    ('fc1 store', 131072, 3072, 768, _lib.EPI_STORE, torch.float16),
    ('fc1 gelu', 131072, 3072, 768, _lib.EPI_BIAS_GELU, torch.float16),
    ('attn_out res', 131072, 768, 768, _lib.EPI_GATE_RESIDUAL, torch.float16),
    ('attn_out store16', 131072, 768, 768, _lib.EPI_STORE, torch.float16),
    ('attn_out accf32', 131072, 768, 768, _lib.EPI_ACCUM_F32, torch.float16),
    ('attn_out wrf32', 131072, 768, 768, _lib.EPI_ACCUM_F32, torch.float16),
    ('fc2 res', 131072, 768, 3072, _lib.EPI_GATE_RESIDUAL, torch.float16),
    ('vocab', 131072, 30522, 768, _lib.EPI_STORE, torch.bfloat16),
  ]
  only = sys.argv[1:] 
  for name, M, N, K, epi, dt in cases:
    if only and not any(o in name for o in only):
      continue
    a = (torch.randn(M, K, device=DEV) * 0.7).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.03).to(dt)
    bias = (torch.randn((N + 255) // 256 * 256, device=DEV) * 0.1).to(dt)
    res = epi in (_lib.EPI_GATE_RESIDUAL, _lib.EPI_ACCUM_F32)
    ldo = (N + 7) // 8 * 8
    out = torch.zeros(M, ldo, dtype=torch.float32 if res else dt, device=DEV)
    gate = (torch.randn(2, N, device=DEV) * 0.1).to(dt)
    cos = torch.randn(L, 32, device=DEV).to(dt)
    sin = torch.randn(L, 32, device=DEV).to(dt)
    stats = torch.empty((N + 127) // 128, M, 2, dtype=torch.float32, device=DEV) if name == 'vocab' else None
    g = _lib.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), bias=bias.data_ptr(),
                      gate=gate.data_ptr() if epi == _lib.EPI_GATE_RESIDUAL else None, rope_cos=cos.data_ptr(), rope_sin=sin.data_ptr(),
                      lda=K, ldw=K, ld_out=ldo, gate_stride=N, rows_per_mod=M // 2, M=M, N=N, K=K,
                      dtype=_lib.DTYPE_CODE[dt], epilogue=epi, seq_len=L, head_dim=64,
                      rope_cols=2 * N // 3 if epi == _lib.EPI_QKV_ROPE else 0,
                      stats=stats.data_ptr() if stats is not None else None, stats_ld=M, stats_mask_col=-1,
                      accumulate=1 if 'accf32' in name else 0)
    sp = _lib.stream_ptr()

    def fn():
      _lib.check(_lib.lib.ddg_gemm(g, sp))
    ms, tf, clk = sustained(fn, 2.0 * M * N * K, h)
    line = f'{name:14s} M{M} N{N} K{K}: ddg {ms:7.3f} ms {tf:7.0f} TF/s @ {clk} MHz'
    if epi == _lib.EPI_STORE and N < 4000:
      wt = w.t()
      ms2, tf2, clk2 = sustained(lambda: torch.matmul(a, wt), 2.0 * M * N * K, h)
      line += f' | cuBLAS {ms2:7.3f} ms {tf2:7.0f} TF/s @ {clk2} MHz'
    print(line, flush=True)
    del a, w, out
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
