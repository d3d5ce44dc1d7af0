#!/bin/bash
# first GPU pass of round 2: full GPU suite, reference-on-GPU fixtures, bench, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -s -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/test_gpu.log
timeout 600 python tests/golden/make_golden.py --gpu --out gpurun_out/golden_gpu > gpurun_out/golden_gpu.log 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 600 python scripts/reference_on_gpu.py > gpurun_out/reference_on_gpu.log 2>&1
tail -5 gpurun_out/test_gpu.log
tail -c 1500 gpurun_out/bench.log
cat gpurun_out/reference_on_gpu.log | grep impl
