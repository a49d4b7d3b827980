#!/bin/bash
# Builds fast.cu variants on the GPU box and times fast_cells over 256 frames.
for v in "-DDSDTM_FAST_UNROLL_A=1" "-DDSDTM_FAST_UNROLL_A=2" "-DDSDTM_FAST_UNROLL_A=3" "-DDSDTM_FAST_UNROLL_A=6"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/fast.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E " error |fast_kernel" -A3 | grep -E "error|Used"
  timeout 100 python scripts/prof_fast.py 3 2>&1 | tail -1
done
