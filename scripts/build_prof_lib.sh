#!/bin/bash
# developer build of the library with extra -D flags -> build/libddg_b200_$1.so (select it with DDG_B200_LIB)
# usage: scripts/build_prof_lib.sh prof -DDDG_LN_PROF
set -e
tag=$1; shift
cd "$(dirname "$0")/../discrete-diffusion-guidance_b200"
mkdir -p ../build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --threads 8 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --shared "$@" \
  csrc/capi.cu csrc/sample_step.cu csrc/sample_fused.cu csrc/sample_group.cu csrc/sample_aux.cu csrc/dit_ops.cu csrc/dit_bwd.cu \
  csrc/attention_bwd.cu csrc/attention_bwd_tc.cu csrc/gemm_tcgen05.cu csrc/attention.cu -o ../build/libddg_b200_$tag.so 2>&1 | grep -v "^$" | grep -iE "error|warning: v" || true
ls -la ../build/libddg_b200_$tag.so
