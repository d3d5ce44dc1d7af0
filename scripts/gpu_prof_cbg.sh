#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -W ignore scripts/profile_cbg.py c5_cbg 512 > gpurun_out/prof_c5_cbg.log 2>&1; head -32 gpurun_out/prof_c5_cbg.log | cut -c1-170
timeout 600 python -W ignore scripts/profile_cbg.py c4 4096 > gpurun_out/prof_c4.log 2>&1; head -30 gpurun_out/prof_c4.log | cut -c1-170
