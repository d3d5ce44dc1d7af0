"""Debug: timing of the fused-LayerNorm residual GEMM at the exact-CBG classifier shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import test_gpu_fused_ln as t
DEV = 'cuda'

def setup(M, N, K, dt):
  g = torch.Generator().manual_seed(1)
  a = torch.randn(M, K, generator=g).to(dt).to(DEV)
  w = (torch.randn(N, K, generator=g) * 0.05).to(dt).to(DEV)
  gate = torch.randn(1, N, generator=g).to(dt).to(DEV)
  lw = torch.ones(N, device=DEV)
  mod = torch.randn(1, 2 * N, generator=g).to(dt).to(DEV)
  x = torch.randn(M, N, generator=g).to(DEV)
  out = torch.zeros(M, N, dtype=dt, device=DEV)
  return a, w, gate, lw, mod, x, out

def launch(s, work, M=None, fused=True):
  a, w, gate, lw, mod, x, out = s
  N = w.shape[0]
  t._residual_gemm(a, w, x, None, gate, N, 1 << 40,
                   ln=(lw, mod[:, :N], mod[:, N:], 2 * N, a.dtype, out) if fused else None, work=work, M=M)

def timeit(fn, n=3):
  res = []
  for _ in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    res.append(round(e0.elapsed_time(e1), 3))
  return res

for (M, N, K) in [(163840, 512, 512), (491520, 512, 512), (491520, 512, 2048), (491520, 768, 768), (131072, 768, 768)]:
  s = setup(M, N, K, torch.bfloat16)
  work = t._work(M, N)
  print(M, N, K, 'fused', timeit(lambda: launch(s, work)), 'unfused', timeit(lambda: launch(s, work, fused=False)),
        'small-M then full', timeit(lambda: (launch(s, work, M=M // 3 // 256 * 256), launch(s, work))))
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    launch(s, work); launch(s, work, M=M // 3 // 256 * 256); launch(s, work)
  print('   graph of 3 launches', timeit(lambda: g.replay()))
  del s, work, g
