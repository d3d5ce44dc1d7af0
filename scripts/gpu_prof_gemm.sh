#!/bin/bash
# usage: gpu_prof_gemm.sh <tag> : ncu --set full of the 5 GEMM launches of one block + the vocab head
TAG=${1:-gemm}
mkdir -p gpurun_out
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 93 -c 5 -f -o gpurun_out/${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
echo "$TAG rc=$?"
