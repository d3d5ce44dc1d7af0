#!/bin/bash
# attention forward: shorter thread-phase latency chains -- parity, then same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dit.py tests/test_gpu_width.py tests/test_gpu_cache.py -m gpu -q -x -p no:cacheprovider -k "attention or forward or dit or width or cache" 2>&1 | tail -3
bash scripts/ab_bench.sh build/libddg_b200_A.so build/libddg_b200_B.so 2
