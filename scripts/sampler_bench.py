"""Micro-benchmark of the fused sampler kernels alone (config-3 shaped rows,
bf16 logits larger than L2).  `python scripts/sampler_bench.py [rows]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch

import helpers
from ddg_b200 import _lib


def main():
  rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
  V, ld, L = 30522, 30528, 128
  dev = torch.device('cuda')
  g = torch.Generator(device=dev).manual_seed(0)
  la = torch.empty(rows, ld, dtype=torch.bfloat16, device=dev)
  lb = torch.empty(rows, ld, dtype=torch.bfloat16, device=dev)
  for t in (la, lb):
    for r0 in range(0, rows, 4096):
      t[r0:r0 + 4096] = (torch.randn(min(4096, rows - r0), ld, device=dev, generator=g) * 2.0).to(torch.bfloat16)
  st_a = torch.cat([helpers.make_stats(la[r0:r0 + 4096], V) for r0 in range(0, rows, 4096)], 1).contiguous()
  st_b = torch.cat([helpers.make_stats(lb[r0:r0 + 4096], V) for r0 in range(0, rows, 4096)], 1).contiguous()
  xt = torch.randint(0, V, (rows,), device=dev, generator=g)
  xs = torch.empty_like(xt)
  mct = torch.tensor([0.5], device=dev)
  mcs = torch.tensor([0.499], device=dev)
  bytes_cfg = rows * (2 * V * 2 + 16)
  for name, kw in (('wide cfg round', dict(race_group=64, guidance=_lib.GUIDE_CFG, round_bf16=1)),
                   ('wide cfg exact', dict(race_group=64, guidance=_lib.GUIDE_CFG, round_bf16=0)),
                   ('wide null round', dict(race_group=64, guidance=_lib.GUIDE_NONE, round_bf16=1)),
                   ('group cfg round', dict(race_group=8, guidance=_lib.GUIDE_CFG, round_bf16=1)),
                   ('group cfg exact', dict(race_group=8, guidance=_lib.GUIDE_CFG, round_bf16=0)),
                   ('group null round', dict(race_group=8, guidance=_lib.GUIDE_NONE, round_bf16=1)),
                   ('column cfg round', dict(race_group=0, guidance=_lib.GUIDE_CFG, round_bf16=1))):
    two = kw['guidance'] == _lib.GUIDE_CFG
    a = _lib.SampleStepArgs(
      logits_a=la.data_ptr(), logits_b=lb.data_ptr() if two else 0, clf=0, noise=0, xt=xt.data_ptr(),
      mc_t=mct.data_ptr(), mc_s=mcs.data_ptr(), step_ptr=0, seed_ptr=0, xs=xs.data_ptr(), q_out=0,
      rows=rows, ld_a=ld, ld_b=ld if two else 0, ld_clf=0, ld_noise=0, ld_q=0, seq_len=L, vocab=V,
      mc_seq_stride=0, mc_step_stride=0, logits_dtype=_lib.BF16, diffusion=_lib.UNIFORM,
      guidance=kw['guidance'], subs=0, mask_logit=0, mask_index=103, round_bf16=kw['round_bf16'],
      select=_lib.SELECT_RACE, fast_math=1, gamma=2.0, seed=1, step=3, stats_chunks=st_a.shape[0],
      stats_a=st_a.data_ptr(), stats_b=st_b.data_ptr() if two else 0, stats_ld=rows,
      race_group=kw['race_group'])
    s = _lib.stream_ptr()
    for _ in range(3):
      _lib.check(_lib.lib.ddg_sample_step(a, s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
      _lib.check(_lib.lib.ddg_sample_step(a, s))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    b = bytes_cfg if two else rows * (V * 2 + 16)
    print(f'{name:18s} {ms:7.3f} ms  {b / ms / 1e6:7.0f} GB/s  ({b / ms / 1e6 / 6446.9 * 100:4.1f}% of 6447)  '
          f'scaled to 65536 rows: {ms * 65536 / rows:.3f} ms')


if __name__ == '__main__':
  main()
