#!/bin/bash
# quick iteration: gemm debug + chosen test files + bench
mkdir -p gpurun_out
timeout -s KILL 300 python -W ignore scripts/gemm_debug.py 2>&1 | grep -E "bad [1-9]|time|Error|error" | head -20
bash scripts/gpu_tests.sh "$@"
timeout -s KILL 900 python -W ignore bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],3), 'clocks', d['clocks'])
    print('gemm roofline', d['roofline']['achieved'], d['roofline']['frac'], '| sampler', d['roofline_sampler']['achieved'], d['roofline_sampler']['frac'], '| step frac', d['step_roofline']['frac'])
    for k, v in d['kernels'].items():
        print(f"  {k:14s} {v['ms_per_step']:8.3f} ms  x{v['launches']:3d}  {v['share']*100:5.1f}%  {v.get('tflops','')}")
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
