#!/bin/bash
# fused-LayerNorm GEMM: unit tests first (short timeout: a protocol bug must not hang the box), then the
# suites that run whole forwards, then an A/B of the bench step (DDG_FUSE_LN=0 -> separate launches)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_ln.py -m gpu -q -x -p no:cacheprovider > gpurun_out/test_fused_ln.log 2>&1
echo "fused_ln exit $?"; tail -15 gpurun_out/test_fused_ln.log
grep -q "passed" gpurun_out/test_fused_ln.log && ! grep -q "failed" gpurun_out/test_fused_ln.log || exit 0
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/test_gpu.log
B="python -W ignore bench.py --steps 60 --warmup 5 --no-configs --no-cpu-baseline"
for i in 1 2; do
  DDG_FUSE_LN=0 $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A unfused', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'])"
  $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B fused  ', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'])"
done
