#!/bin/bash
# ncu evidence for profiles/: launch list of one eager step + full sections of
# the top kernels.  Plain run first (must exit 0), then ncu (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_plain.log; exit 1; }
cat gpurun_out/profile_plain.log
# 2 steps x 91 launches; skip model-load casts (-s via kernel regex instead)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_kernel|attention_kernel|ln_modulate_kernel|sample_step_kernel|cond_embed_kernel|adaln_kernel|step_advance_kernel' -c 182 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/ncu_launches.log
ncu --set full --clock-control none --import-source on -k regex:sample_step_kernel -s 1 -c 1 -f -o gpurun_out/prof_sampler $CMD > gpurun_out/ncu_sampler.log 2>&1
echo "sampler rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 61 -c 5 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 12 -c 1 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out/
