#!/bin/bash
# 16-bit GEMM epilogue with the next chunk's tensor-memory load in flight: parity, burst timings, A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dit.py tests/test_gpu_width.py tests/test_gpu_fused_ln.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for l in A B; do echo "== $l"; DDG_B200_LIB=$PWD/build/libddg_b200_$l.so SUSTAIN_S=1.0 timeout 300 python -W ignore scripts/gemm_sustained.py "qkv rope" "fc1 gelu" "fc1 store" vocab 2>&1 | tail -5; done
bash scripts/ab_bench.sh build/libddg_b200_A.so build/libddg_b200_B.so 2
