#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/test_gpu.log
tail -8 gpurun_out/test_gpu.log
timeout 600 python -W ignore bench.py --config c2 --steps 50 --warmup 5 2>gpurun_out/c2.err | tee gpurun_out/bench_c2.log | cut -c1-600
B="python -W ignore bench.py --steps 60 --warmup 5 --no-configs --no-cpu-baseline"
for i in 1 2; do
  DDG_B200_LIB=$PWD/build/libddg_b200_A.so $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A', d['ms_per_step'], d['kernels']['gemm_vocab'], d['clocks']['sm_mhz'])"
  $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', d['ms_per_step'], d['kernels']['gemm_vocab'], d['clocks']['sm_mhz'])"
done
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 97 -c 1 -f -o gpurun_out/r02_gemm_vocab $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/r02_gemm_vocab.ncu-rep gpurun_out/r02_gemm_vocab_summary.csv
