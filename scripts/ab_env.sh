#!/bin/bash
# Same-box A/B of one build under two settings of an environment switch.
# usage: ab_env.sh VAR valueA valueB [rounds]     e.g.  ab_env.sh DDG_FUSE_LN 0 1 2
mkdir -p gpurun_out
VAR=$1; A=$2; B=$3; R=${4:-2}
for r in $(seq 1 $R); do
  for val in $A $B; do
    env $VAR=$val timeout -s KILL 600 python -W ignore bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ab.log 2>&1
    python - <<PY
import json
d = json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
k = d['kernels']
print('$VAR=$val', 'ms/step %.3f' % d['ms_per_step'], 'e2e %.2f' % d['e2e']['value'], 'clk', d['clocks']['sm_mhz'], ' '.join(f"{n.replace('gemm_','')}={v['ms_per_step']:.2f}" for n, v in k.items()))
PY
  done
done
