#!/bin/bash
# Same-box A/B of two builds of libddg_b200.so (boxes differ by +-8 % under the power cap).
# usage: ab_bench.sh <libA.so> <libB.so> [rounds]
mkdir -p gpurun_out
A=$1; B=$2; R=${3:-2}
for r in $(seq 1 $R); do
  for L in $A $B; do
    DDG_B200_LIB=$PWD/$L timeout -s KILL 600 python -W ignore bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/ab.log 2>&1
    python - <<PY
import json
d = json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
k = d['kernels']
print('$L', 'ms/step %.3f' % d['ms_per_step'], 'clk', d['clocks']['sm_mhz'], ' '.join(f"{n.replace('gemm_','')}={v['ms_per_step']:.2f}" for n, v in k.items()))
PY
  done
done
