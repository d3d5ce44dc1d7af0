#!/bin/bash
# second GPU pass: full suite again, vocab-GEMM L2 policy A/B (A = previous build), D-CBG profiles,
# ncu full capture of the vocab GEMM (dram bytes)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/test_gpu.log
tail -4 gpurun_out/test_gpu.log
B="python -W ignore bench.py --steps 60 --warmup 5 --no-configs --no-cpu-baseline"
for i in 1 2; do
  DDG_B200_LIB=$PWD/build/libddg_b200_A.so $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A', d['ms_per_step'], d['kernels']['gemm_vocab'], d['clocks']['sm_mhz'])"
  $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', d['ms_per_step'], d['kernels']['gemm_vocab'], d['clocks']['sm_mhz'])"
done
timeout 600 python -W ignore scripts/profile_cbg_amazon.py 1024 > gpurun_out/profile_cbg_amazon.log 2>&1; tail -28 gpurun_out/profile_cbg_amazon.log
timeout 600 python -W ignore scripts/profile_cbg.py > gpurun_out/profile_cbg.log 2>&1; tail -28 gpurun_out/profile_cbg.log
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 97 -c 1 -f -o gpurun_out/r02_gemm_vocab $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gemm.log
python scripts/ncu_summary.py gpurun_out/r02_gemm_vocab.ncu-rep gpurun_out/r02_gemm_vocab_summary.csv
