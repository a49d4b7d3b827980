#!/bin/bash
# Shared-memory carveout preference of sparse_align_kernel<4> (percent of the 256 KB L1/shared array), 4096-pair step.
for v in "" "-DDSDTM_SA_CARVEOUT=45" "-DDSDTM_SA_CARVEOUT=60" "-DDSDTM_SA_CARVEOUT=100"; do
  echo "=== default $v"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E " error " -A3
  timeout 120 python scripts/prof_step.py --pairs 4096 --steps 5 --direct 2>&1 | tail -2 | head -1 | cut -c60-110
done
