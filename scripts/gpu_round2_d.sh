#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dit.py tests/test_gpu_width.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 900 python -W ignore bench.py --steps 60 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.log
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 97 -c 1 -f -o gpurun_out/r02_gemm_vocab $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/r02_gemm_vocab.ncu-rep gpurun_out/r02_gemm_vocab_summary.csv
