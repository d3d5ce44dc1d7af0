#!/bin/bash
# One gpurun call: diagnostics + GPU test files (each in its own process, each
# under its own timeout so a hung kernel cannot eat the box) + smoke + bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gemm_debug"; timeout -s KILL 300 python -W ignore scripts/gemm_debug.py > gpurun_out/gemm_debug.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/gemm_debug.log
for t in test_gpu_dit test_gpu_sampler test_gpu_e2e; do
  echo "== $t"
  timeout -s KILL 1200 python -W ignore -m pytest tests/$t.py -q -m gpu --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "rc=$?"; tail -25 gpurun_out/$t.log
done
echo "== smoke"; timeout -s KILL 600 python -W ignore -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench"; timeout -s KILL 900 python -W ignore bench.py --steps 30 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/bench.log
