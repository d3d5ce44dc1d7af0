#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -W ignore scripts/profile_cbg.py c4_exact 64 > gpurun_out/dbg_exact_fused.log 2>&1
DDG_FUSE_LN=0 timeout 600 python -W ignore scripts/profile_cbg.py c4_exact 64 > gpurun_out/dbg_exact_unfused.log 2>&1
head -12 gpurun_out/dbg_exact_fused.log | cut -c1-200; echo; head -12 gpurun_out/dbg_exact_unfused.log | cut -c1-200
