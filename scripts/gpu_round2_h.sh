#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_ln.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/test_gpu.log
for c in c4 c4_exact c5_cbg; do timeout 600 python -W ignore bench.py --config $c --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:40], d['ms_per_step'], d['value'])"; done
