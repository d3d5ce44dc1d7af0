#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full sections of the
# top kernels.  Plain runs first (must exit 0), then ncu (B200_PROFILING.md).
# usage: gpu_profile.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
KRE='gemm_kernel|attention_kernel|attention_tc_kernel|ln_modulate_kernel|ln_modulate_stream_kernel|sample_step_kernel|sample_fused_kernel|sample_group_kernel|cond_embed_kernel|adaln_kernel|step_advance_kernel'
# (1) launch list of bench.py itself: 3 steps' worth of our kernels (91 launches per step)
BENCH="python -W ignore bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/${TAG}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -20 gpurun_out/${TAG}_bench_plain.log; exit 1; }
tail -c 600 gpurun_out/${TAG}_bench_plain.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KRE" -c 273 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_launches.log
# (2) full sections from eager steps of the same workload
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_plain.log; exit 1; }
cat gpurun_out/profile_plain.log
# step 2, block 11: qkv, attn_out, fc1, fc2, then the vocab head (gemm launches 93..97)
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 93 -c 5 -f -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'sample_fused_kernel|sample_step_kernel|sample_group_kernel' -s 1 -c 1 -f -o gpurun_out/${TAG}_sampler $CMD > gpurun_out/ncu_sampler.log 2>&1
echo "sampler rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'attention_kernel|attention_tc_kernel' -s 12 -c 1 -f -o gpurun_out/${TAG}_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'ln_modulate_kernel|ln_modulate_stream_kernel' -s 1 -c 1 -f -o gpurun_out/${TAG}_ln $CMD > gpurun_out/ncu_ln.log 2>&1
echo "ln rc=$?"
rm -f gpurun_out/${TAG}_ncu_full_summary.csv
for f in gemm sampler attn ln; do python scripts/ncu_summary.py gpurun_out/${TAG}_$f.ncu-rep gpurun_out/${TAG}_ncu_full_summary.csv; done
ls -la gpurun_out/ | head -30
