#!/bin/bash
# GPU test files, each in its own process and under its own timeout.
mkdir -p gpurun_out
for t in ${@:-test_gpu_dit test_gpu_sampler test_gpu_e2e}; do
  echo "== $t"
  timeout -s KILL 1200 python -W ignore -m pytest tests/$t.py -q -m gpu --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "rc=$?"; tail -12 gpurun_out/$t.log
done
