"""One GEMM shape repeated (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddg_b200
from ddg_b200 import _lib
M, N, K = [int(x) for x in sys.argv[1:4]]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
a = torch.randn(M, K, device='cuda').half()
w = (torch.randn(N, K, device='cuda') * 0.05).half()
out = torch.empty(M, (N + 7) // 8 * 8, dtype=torch.float16, device='cuda')
g = _lib.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), lda=K, ldw=K, ld_out=out.stride(0),
                  M=M, N=N, K=K, dtype=_lib.F16, epilogue=_lib.EPI_STORE, rows_per_mod=1, seq_len=1, head_dim=64)
for _ in range(reps):
  _lib.check(_lib.lib.ddg_gemm(g, _lib.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
  _lib.check(_lib.lib.ddg_gemm(g, _lib.stream_ptr()))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'M{M} N{N} K{K}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF/s')
