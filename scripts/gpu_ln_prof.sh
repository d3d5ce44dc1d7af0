#!/bin/bash
# ncu full capture (with source) of the fused-LayerNorm residual GEMMs of the bench step: block 11 of eager step 2
mkdir -p gpurun_out
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 || { tail -20 gpurun_out/profile_plain.log; exit 1; }
cat gpurun_out/profile_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 94 -c 3 -f -o gpurun_out/r02_gemm_ln $CMD > gpurun_out/ncu_gemm_ln.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gemm_ln.log
python scripts/ncu_summary.py gpurun_out/r02_gemm_ln.ncu-rep
