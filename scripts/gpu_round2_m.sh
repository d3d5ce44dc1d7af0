#!/bin/bash
# attention backward alone (tcgen05 vs mma.sync) + one ncu --set full capture of the tcgen05 kernel
mkdir -p gpurun_out
timeout 120 python -W ignore scripts/attn_bwd_bench.py
DDG_ATT_BWD=mma timeout 120 python -W ignore scripts/attn_bwd_bench.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02c_attn_bwd python -W ignore scripts/attn_bwd_bench.py --once > gpurun_out/ncu_attn_bwd.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn_bwd.log
python scripts/ncu_summary.py gpurun_out/r02c_attn_bwd.ncu-rep gpurun_out/r02c_attn_bwd_summary.csv
