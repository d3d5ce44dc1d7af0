#!/bin/bash
# fused-LayerNorm epilogue with TMA stores: parity under a short timeout, phase profile, A/B
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_fused_ln.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "fused ln rc=$?"
timeout -s KILL 600 python -m pytest tests/test_gpu_dit.py tests/test_gpu_width.py tests/test_gpu_cache.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
DDG_B200_LIB=$PWD/build/libddg_b200_prof.so timeout 300 python -W ignore scripts/ln_epilogue_prof.py 2>&1 | tail -4
bash scripts/ab_bench.sh build/libddg_b200_A.so build/libddg_b200_B.so 2
