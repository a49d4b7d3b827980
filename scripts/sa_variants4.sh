#!/bin/bash
# round 2: 4 warps per pair compiled for FOUR resident CTAs per SM (128 registers): 16 warps per SM, 3 rounds per pass;
# batch sizes around one wave (configs[4] partitioned over 4 / 8 GPUs leaves 1024 / 512 pairs per GPU)
set -e
for v in "-DDSDTM_SA_MINB4=3" "-DDSDTM_SA_MINB4=4"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E "sparse_align_kernelILi4E" -A3 | grep -E "Used|spill" | head -2
  for n in 296 512 592 1024 2048 4096; do
    echo "--- pairs $n"
    timeout 300 python scripts/sa_sweep.py --pairs $n --combos 0:3,0:4,0:5 2>&1 | tail -3 | cut -c1-75
  done
done
