#!/bin/bash
# N = 8 (and 4): the driver's launch line for the scaling run
mkdir -p gpurun_out
for N in ${NS:-8 4}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r02d_bench_n$N.log 2> gpurun_out/r02d_bench_n$N.err
echo "n$N rc=$?"; grep -iE "error|Traceback" gpurun_out/r02d_bench_n$N.err | head -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench_n$N.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','steps','warmup')}, d['config']['batch_per_gpu'], d['config']['global_batch'])
print('weak', d.get('weak')); print('e2e', d['e2e']['value']); print('clocks', d['clocks'])
for k,v in d.get('configs',{}).items(): print(k,{a:b for a,b in v.items() if a!='workload'})
PY
done
