"""Phase breakdown of the fused-LayerNorm GEMM epilogue (developer build with -DDDG_LN_PROF:
`scripts/build_prof_lib.sh`, selected through DDG_B200_LIB).  For the two residual shapes of
config 3 (attn-out K=768, fc2 K=3072): sustained ms per launch and, for one epilogue warp of one
CTA, the average clock64() cycles per tile spent waiting for the accumulator, in pass 1, in the
residual store pass, polling the neighbours' partials and in pass 2; plus what the MMA thread
waited for a free accumulator.  `DDG_B200_LIB=$PWD/build/libddg_b200_prof.so python scripts/ln_epilogue_prof.py`"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ddg_b200
from ddg_b200 import _lib

DEV = 'cuda'
lib = _lib.lib
has_prof = hasattr(lib, 'ddg_debug_ln_prof')


def run(name, M, N, K, fused, iters=300):
  dt = torch.float16
  a = (torch.randn(M, K, device=DEV) * 0.7).to(dt)
  w = (torch.randn(N, K, device=DEV) * 0.03).to(dt)
  bias = (torch.randn(N, device=DEV) * 0.1).to(dt)
  x = torch.randn(M, N, device=DEV)
  gate = (torch.randn(2, N, device=DEV) * 0.1).to(dt)
  lnw = torch.ones(N, device=DEV)
  mod = (torch.randn(2, 2 * N, device=DEV) * 0.1).to(dt)
  out16 = torch.empty(M, N, dtype=dt, device=DEV)
  work = torch.zeros(int(lib.ddg_gemm_ln_work_bytes(M, N)), dtype=torch.uint8, device=DEV)
  g = _lib.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=x.data_ptr(), bias=bias.data_ptr(), gate=gate.data_ptr(),
                    lda=K, ldw=K, ld_out=N, gate_stride=N, rows_per_mod=M // 2, M=M, N=N, K=K,
                    dtype=_lib.DTYPE_CODE[dt], epilogue=_lib.EPI_GATE_RESIDUAL, seq_len=128, head_dim=64)
  if fused:
    g.ln_weight, g.ln_shift, g.ln_scale = lnw.data_ptr(), mod.data_ptr(), mod.data_ptr() + 2 * N
    g.ln_mod_stride, g.ln_dtype = 2 * N, _lib.DTYPE_CODE[dt]
    g.ln_out, g.ln_work = out16.data_ptr(), work.data_ptr()
  sp = _lib.stream_ptr()
  for _ in range(iters // 2):
    _lib.check(lib.ddg_gemm(g, sp))
  torch.cuda.synchronize()
  buf = (C.c_ulonglong * 8)()
  if has_prof:
    lib.ddg_debug_ln_prof(buf, 1)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    _lib.check(lib.ddg_gemm(g, sp))
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / iters
  line = f'{name:22s} {ms:7.4f} ms/launch'
  if has_prof and fused:
    lib.ddg_debug_ln_prof(buf, 1)
    v = [int(b) for b in buf]
    tiles = max(v[5], 1)
    names = ['wait_acc', 'pass1', 'store', 'poll', 'pass2']
    line += f' | tiles/launch {tiles / iters:.1f} | cycles/tile: ' + ' '.join(
      f'{nm} {v[i] / tiles:7.0f}' for i, nm in enumerate(names))
    line += f' | sum {sum(v[:5]) / tiles:7.0f} | mma_wait_acc {v[6] / tiles:7.0f}'
  print(line, flush=True)


if __name__ == '__main__':
  M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
  for _ in range(2):
    run('attn_out K=768 +LN', M, 768, 768, True)
    run('attn_out K=768', M, 768, 768, False)
    run('fc2 K=3072 +LN', M, 768, 3072, True)
    run('fc2 K=3072', M, 768, 3072, False)
