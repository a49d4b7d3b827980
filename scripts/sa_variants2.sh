#!/bin/bash
# round 2: register caps of the 3- and 5-warp kernels, built on the GPU box
set -e
for v in "-DDSDTM_SA_MINB3=4" "-DDSDTM_SA_MINB3=5" "-DDSDTM_SE3_SERIES=0"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E "sparse_align_kernelILi3E" -A3 | grep -E "Used|spill" | head -2
  timeout 300 python scripts/sa_sweep.py --combos 0:3,0:4,0:5 2>&1 | tail -3 | cut -c1-75
done
