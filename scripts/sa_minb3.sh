#!/bin/bash
# CTAs per SM the 3-warp sparse-alignment kernel is compiled for (register cap 65536 / (96 * N)), 4096-pair step.
for v in "-DDSDTM_SA_MINB3=4" "-DDSDTM_SA_MINB3=5" "-DDSDTM_SA_MINB3=6"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E " error " -A3
  grep -n "sparse_align_kernelILi3E" -A3 dsdtm_b200/lib/ptxas.log | grep -E "Used|spill" | head -2
  timeout 120 python scripts/prof_step.py --pairs 4096 --steps 10 --direct 2>&1 | tail -2 | head -1 | cut -c60-112
done
