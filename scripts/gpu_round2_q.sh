#!/bin/bash
# end-of-round check: what the driver runs (GPU suite, smoke, default bench) + the bench launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/test_gpu.log
timeout 600 python -W ignore -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -W ignore bench.py > gpurun_out/r02d_bench.log 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, d['e2e']['value'], d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline'].get('gemms_without_fused_layernorm'), 'sampler', d['roofline_sampler']['frac'], d['roofline_sampler']['alone']['frac'])
for k,v in d['configs'].items(): print(k, v.get('ms_per_step'), v.get('value'))
PY
KRE='gemm_kernel|attention_kernel|attention_tc_kernel|ln_modulate_kernel|ln_modulate_stream_kernel|sample_step_kernel|sample_fused_kernel|sample_group_kernel|cond_embed_kernel|adaln_kernel|step_advance_kernel|seq_compact_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KRE" -c 268 --csv --log-file gpurun_out/r02d_launches.csv python -W ignore bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
