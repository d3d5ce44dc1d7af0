"""Attention input-gradient kernels alone (first-order D-CBG classifier backward): the tcgen05 /
tensor-memory kernel (attention_bwd_tc.cu) and, with DDG_ATT_BWD=mma, the mma.sync one, at the
C5 shape (512 sequences x 128 tokens x 12 heads, bf16) and the C4 shape (4096 x 32 x 8).
`python scripts/attn_bwd_bench.py`; under ncu: `--once` runs each shape once after a warm-up."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ddg_b200
from ddg_b200 import _lib

DEV = 'cuda'
once = '--once' in sys.argv
for B, L, H in ((512, 128, 12), (4096, 32, 8)):
  dt = torch.bfloat16
  qkv = (torch.randn(B * L, 3 * H * 64, device=DEV) * 1.2).to(dt)
  go = torch.randn(B * L, H * 64, device=DEV).to(dt)
  cos = torch.rand(L, 32, device=DEV).to(dt)
  sin = torch.rand(L, 32, device=DEV).to(dt)
  out = torch.empty_like(qkv)

  def fn():
    _lib.check(_lib.lib.ddg_attention_backward(
      qkv.data_ptr(), go.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, L, H, 64, 0,
      _lib.DTYPE_CODE[dt], out.data_ptr(), _lib.stream_ptr()))
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  n = 1 if once else 50
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n):
    fn()
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  # algorithmic bytes: q, k, v, dO read once, dq, dk, dv written once (16-bit)
  nbytes = B * L * H * 64 * 2 * 7
  # tensor work: S, dP, S^T, dP^T (K = 64) and dq, dk, dv (K = L): per (seq, head)
  flops = B * H * (4 * 2 * L * L * 64 + 3 * 2 * L * L * 64)
  print(f'attention_backward B={B} L={L} H={H} [{os.environ.get("DDG_ATT_BWD", "tc")}]: {ms:.4f} ms  '
        f'{nbytes / ms / 1e6:.0f} GB/s algorithmic  {flops / ms / 1e9:.0f} TFLOP/s', flush=True)
