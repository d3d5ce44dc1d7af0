#!/bin/bash
# fused-LayerNorm epilogue v2 (TMA-fed residual rows): parity, then the phase profile
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_ln.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
echo "pytest exit $?"
DDG_B200_LIB=$PWD/build/libddg_b200_prof.so timeout 300 python -W ignore scripts/ln_epilogue_prof.py 2>&1 | tail -12
