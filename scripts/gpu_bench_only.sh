#!/bin/bash
mkdir -p gpurun_out
for seg in "$@"; do
DDG_SEG_SEQS=$seg timeout -s KILL 900 python -W ignore bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$seg.log 2>&1; echo "seg=$seg rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_$seg.log').read().strip().splitlines()[-1])
    print('  value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],3), 'clk', d['clocks']['sm_mhz'])
    print('  ' + ' | '.join(f"{k} {v['ms_per_step']:.2f}" for k, v in d['kernels'].items()))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_$seg.log').read()[-1500:])
PY
done
