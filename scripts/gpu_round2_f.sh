#!/bin/bash
# re-entry pass of round 2: full GPU suite, bench (both arms), ncu launch list + full sections (tag r02)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/test_gpu.log
tail -5 gpurun_out/test_gpu.log
timeout 400 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
cut -c1-300 gpurun_out/bench_ref.log
timeout 1200 python -W ignore bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.log
bash scripts/gpu_profile.sh r02 2>&1 | tail -40
