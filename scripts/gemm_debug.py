"""Diagnostic for the tcgen05 GEMM: structured inputs that expose descriptor /
swizzle / layout mistakes, printed compactly.  Not a test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddg_b200
from ddg_b200 import _lib

DEV = 'cuda'


def run(a, w, N=None, epi=_lib.EPI_STORE, dtype=torch.float16):
  M, K = a.shape
  N = w.shape[0]
  out = torch.full((M, (N + 7) // 8 * 8), -777.0, dtype=dtype, device=DEV)
  g = _lib.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), lda=a.stride(0),
                    ldw=w.stride(0), ld_out=out.stride(0), M=M, N=N, K=K,
                    dtype=_lib.DTYPE_CODE[dtype], epilogue=epi, rows_per_mod=1, seq_len=1,
                    head_dim=64)
  _lib.check(_lib.lib.ddg_gemm(g, _lib.stream_ptr()))
  torch.cuda.synchronize()
  return out[:, :N].float()


def report(name, got, ref):
  err = (got - ref).abs()
  bad = err > 1e-2 + 1e-2 * ref.abs()
  print(f'[{name}] shape {tuple(got.shape)} max_err {err.max().item():.4g} '
        f'bad {int(bad.sum())}/{bad.numel()} untouched {int((got == -777).sum())}')
  if bad.any():
    idx = bad.nonzero()[:6].tolist()
    for (i, j) in idx:
      print(f'    ({i},{j}) got {got[i, j].item():.4f} ref {ref[i, j].item():.4f}')
    rows_bad = bad.any(1).nonzero().flatten()
    cols_bad = bad.any(0).nonzero().flatten()
    print(f'    bad rows: n={rows_bad.numel()} first {rows_bad[:8].tolist()} last {rows_bad[-4:].tolist()}')
    print(f'    bad cols: n={cols_bad.numel()} first {cols_bad[:8].tolist()} last {cols_bad[-4:].tolist()}')


def main():
  torch.manual_seed(0)
  for (M, N, K) in [(128, 256, 16), (128, 256, 64), (128, 256, 128), (128, 256, 768),
                    (256, 512, 256), (100, 200, 64), (4096, 2304, 768)]:
    a = (torch.randn(M, K) * 0.5).half().to(DEV)
    w = (torch.randn(N, K) * 0.1).half().to(DEV)
    got = run(a, w)
    report(f'rand M{M} N{N} K{K}', got, a.float() @ w.float().t())
  # selector pattern: A picks column k = m % K of W  ->  out[m, n] = W[n, m % K]
  M, N, K = 128, 256, 64
  a = torch.zeros(M, K)
  a[torch.arange(M), torch.arange(M) % K] = 1.0
  w = (torch.arange(N)[:, None] + torch.arange(K)[None, :] / 100.0)
  got = run(a.half().to(DEV), w.half().to(DEV))
  ref = w.half().float().to(DEV)[:, torch.arange(M) % K].t()
  report('selector', got, ref)
  print('selector row 1 got[:8]:', got[1, :8].tolist())
  print('selector row 1 ref[:8]:', ref[1, :8].tolist())
  print('selector col 0 got[:8]:', got[:8, 0].tolist())
  print('selector col 0 ref[:8]:', ref[:8, 0].tolist())
  # bf16
  a = (torch.randn(300, 128) * 0.5).bfloat16().to(DEV)
  w = (torch.randn(384, 128) * 0.1).bfloat16().to(DEV)
  report('bf16', run(a, w, dtype=torch.bfloat16), a.float() @ w.float().t())
  # timing of the big shapes
  for (M, N, K) in [(131072, 2304, 768), (131072, 768, 768), (131072, 3072, 768),
                    (131072, 768, 3072), (16384, 30522, 768)]:
    a = torch.randn(M, K, device=DEV).half()
    w = (torch.randn(N, K, device=DEV) * 0.05).half()
    out = torch.empty(M, (N + 7) // 8 * 8, dtype=torch.float16, device=DEV)
    g = _lib.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), lda=K, ldw=K,
                      ld_out=out.stride(0), M=M, N=N, K=K, dtype=_lib.F16,
                      epilogue=_lib.EPI_STORE, rows_per_mod=1, seq_len=1, head_dim=64)
    for _ in range(3):
      _lib.check(_lib.lib.ddg_gemm(g, _lib.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
      _lib.check(_lib.lib.ddg_gemm(g, _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tf = 2 * M * N * K / ms / 1e9
    e0.record()
    for _ in range(10):
      torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f'[time] M{M} N{N} K{K}: ddg {ms:.3f} ms {tf:.0f} TF/s | cuBLAS {ms2:.3f} ms '
          f'{2 * M * N * K / ms2 / 1e9:.0f} TF/s')


if __name__ == '__main__':
  main()
