#!/bin/bash
# N = 2: the driver's launch line for both arms
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r02c_bench_n2.log 2> gpurun_out/r02c_bench_n2.err
echo "n2 rc=$?"; tail -c 300 gpurun_out/r02c_bench_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02c_bench_n2.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','steps','warmup','gpu_launches')}, d['config']['batch_per_gpu'], d['config']['global_batch'])
print('weak', d.get('weak')); print('e2e', d['e2e']); print('clocks', d['clocks'])
for k,v in d.get('configs',{}).items(): print(k,{a:b for a,b in v.items() if a!='workload'})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
