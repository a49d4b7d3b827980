#!/bin/bash
# round 2: how the first iteration's second moments are accumulated (DSDTM_SA_MOM), built on the GPU box
set -e
for v in "-DDSDTM_SA_MOM=0" "-DDSDTM_SA_MOM=1" "-DDSDTM_SA_MOM=2"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E "sparse_align_kernelILi3E" -A3 | grep -E "Used|spill" | head -2
  timeout 300 python scripts/sa_sweep.py --combos 0:3,0:4 2>&1 | tail -2 | cut -c1-75
done
