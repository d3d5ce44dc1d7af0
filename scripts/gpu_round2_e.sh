#!/bin/bash
# 2-GPU pass: the driver's own launch line for N=2 (strong scaling + weak companion + side configs),
# the reference arm under torchrun, and the full GPU suite on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; tail -3 gpurun_out/test_gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; tail -c 3000 gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-400
