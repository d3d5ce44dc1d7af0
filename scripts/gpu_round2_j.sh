#!/bin/bash
# default bench line at HEAD + reference arm + compute-sanitizer probe on the small-shape kernel tests
mkdir -p gpurun_out
timeout 900 python -W ignore bench.py > gpurun_out/r02c_bench.log 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02c_bench.err
timeout 600 python -W ignore bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.log 2> gpurun_out/r02c_bench_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/r02c_bench_ref.log
timeout 300 compute-sanitizer --tool memcheck python -W ignore -m pytest tests/test_gpu_dit.py -m gpu -q -x -p no:cacheprovider -k "attention_backward_tensor_memory_kernel and 32-4" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/memcheck.log
