#!/bin/bash
# TMA-store epilogue (STORE / QKV_ROPE / BIAS_GELU): parity under a short timeout first, then sustained + A/B
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_dit.py -m gpu -q -x -p no:cacheprovider -k "gemm" 2>&1 | tail -4
echo "gemm tests rc=$?"
timeout -s KILL 600 python -m pytest tests/test_gpu_dit.py tests/test_gpu_width.py tests/test_gpu_cache.py tests/test_gpu_fused_ln.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout -s KILL 400 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x -p no:cacheprovider -k "ar_ or full_size or graph" 2>&1 | tail -3
for l in A B; do echo "== $l"; DDG_B200_LIB=$PWD/build/libddg_b200_$l.so SUSTAIN_S=1.5 timeout 300 python -W ignore scripts/gemm_sustained.py "qkv store" "qkv rope" "fc1 store" "fc1 gelu" vocab 2>&1 | tail -5; done
bash scripts/ab_bench.sh build/libddg_b200_A.so build/libddg_b200_B.so 2
