#!/bin/bash
# round 2: how the first iteration's second moments are accumulated (DSDTM_SA_MOM = 0 select per pixel, 1 always, 2 uniform branch per
# patch row, 3 two instantiations of the feature pass), built on the GPU box. MOMS="1 3" bash scripts/sa_variants3.sh
set -e
for m in ${MOMS:-0 1 2 3}; do
  echo "=== -DDSDTM_SA_MOM=$m"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="-DDSDTM_SA_MOM=$m" python dsdtm_b200/build.py 2>&1 | grep -E "sparse_align_kernelILi3E" -A3 | grep -E "Used|spill" | head -2
  timeout 300 python scripts/sa_sweep.py --combos 0:3,0:4 2>&1 | tail -2 | cut -c1-75
done
