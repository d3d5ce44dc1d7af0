#!/bin/bash
# after the fused-LayerNorm GEMM + captured D-CBG loop: full bench line (all side configs), ncu evidence (tag r02b)
mkdir -p gpurun_out
timeout 1500 python -W ignore bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 5000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
bash scripts/gpu_profile.sh r02b 2>&1 | tail -25
