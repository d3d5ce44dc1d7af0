#!/bin/bash
# ABI v6 (pre-activation side output + GELU_BWD epilogue): parity, then the D-CBG profiles / bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dit.py -m gpu -q -x -p no:cacheprovider -k "gelu or first_order or gemm" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_reference.py tests/test_checkpoint.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -W ignore scripts/profile_cbg.py c5_cbg 512 > gpurun_out/prof_c5_cbg.log 2>&1; head -14 gpurun_out/prof_c5_cbg.log | cut -c1-150
timeout 600 python -W ignore scripts/profile_cbg.py c4 4096 > gpurun_out/prof_c4.log 2>&1; head -14 gpurun_out/prof_c4.log | cut -c1-150
