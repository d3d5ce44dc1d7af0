#!/bin/bash
# end-of-round ncu --set full captures of the kernels that changed late in the round
mkdir -p gpurun_out
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/profile_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_tc_kernel' -s 12 -c 1 -f -o gpurun_out/r02d_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
CBG="python -W ignore scripts/profile_cbg.py c5_cbg 256"
$CBG > gpurun_out/prof_c5_cbg_256.log 2>&1 || { echo "plain cbg failed"; tail -5 gpurun_out/prof_c5_cbg_256.log; exit 1; }
head -3 gpurun_out/prof_c5_cbg_256.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sample_fused_kernel' -s 2 -c 1 -f -o gpurun_out/r02d_sampler_cbg $CBG > gpurun_out/ncu_scbg.log 2>&1
echo "sampler cbg rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel<__nv_bfloat16, 5>|gemm_kernel<__nv_bfloat16, \(int\)5>' -s 13 -c 1 -f -o gpurun_out/r02d_gemm_gelu_bwd $CBG > gpurun_out/ncu_gbwd.log 2>&1
echo "gelu_bwd rc=$?"; tail -2 gpurun_out/ncu_gbwd.log
rm -f gpurun_out/r02d_ncu_full_summary.csv
for f in attn sampler_cbg gemm_gelu_bwd; do [ -f gpurun_out/r02d_$f.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/r02d_$f.ncu-rep gpurun_out/r02d_ncu_full_summary.csv; done
