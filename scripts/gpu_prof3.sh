#!/bin/bash
# usage: gpu_prof3.sh <kernel-regex> <skip> <name>
mkdir -p gpurun_out
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/prof_$3 $CMD > gpurun_out/ncu_$3.log 2>&1
echo "$3 rc=$?"
