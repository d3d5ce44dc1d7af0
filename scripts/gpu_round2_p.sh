#!/bin/bash
# single-pass sampler with the D-CBG term: parity, then the D-CBG profile / bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -q -x -p no:cacheprovider -k "cbg or fused" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_reference.py tests/test_gpu_dit.py tests/test_checkpoint.py -m gpu -q -x -p no:cacheprovider -k "cbg or first_order or approx or checkpoint" 2>&1 | tail -3
timeout 600 python -W ignore scripts/profile_cbg.py c5_cbg 512 > gpurun_out/prof_c5_cbg.log 2>&1; head -8 gpurun_out/prof_c5_cbg.log | cut -c1-150
for c in c4 c5_cbg; do timeout 600 python -W ignore bench.py --config $c --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:40], d['ms_per_step'], d['value'])"; done
