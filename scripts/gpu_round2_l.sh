#!/bin/bash
# programmatic dependent launch: full GPU suite, then same-box A/B of DDG_PDL=0/1 at B=512, B=64 and C1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/test_gpu.log
ab() {  # label, bench args
  for r in 1 2; do for v in 0 1; do
    DDG_PDL=$v timeout 600 python -W ignore bench.py --no-cpu-baseline --no-configs $2 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 PDL=$v ms/step %.4f value %.2f e2e %.2f clk %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))"
  done; done
}
ab B512 "--steps 60 --warmup 5"
ab B64 "--batch 64 --steps 300 --warmup 10"
ab C1 "--config c1 --steps 120 --warmup 8"
