#!/bin/bash
# tcgen05 attention backward: parity (new kernel and the mma.sync one), then the D-CBG profiles
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dit.py -m gpu -q -x -p no:cacheprovider -k "attention_backward" 2>&1 | tail -15
echo "pytest exit $?"
DDG_ATT_BWD=mma timeout 300 python -m pytest tests/test_gpu_dit.py -m gpu -q -x -p no:cacheprovider -k "attention_backward" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_dit.py tests/test_gpu_e2e.py -m gpu -q -x -p no:cacheprovider -k "first_order or cbg or approx" 2>&1 | tail -5
timeout 600 python -W ignore scripts/profile_cbg.py c5_cbg 512 > gpurun_out/prof_c5_cbg.log 2>&1; head -12 gpurun_out/prof_c5_cbg.log | cut -c1-170
timeout 600 python -W ignore scripts/profile_cbg.py c4 4096 > gpurun_out/prof_c4.log 2>&1; head -12 gpurun_out/prof_c4.log | cut -c1-170
