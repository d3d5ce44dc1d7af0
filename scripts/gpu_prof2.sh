#!/bin/bash
mkdir -p gpurun_out
CMD="python -W ignore scripts/profile_step.py --batch 512 --steps 2"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_plain.log; exit 1; }
cat gpurun_out/profile_plain.log
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
