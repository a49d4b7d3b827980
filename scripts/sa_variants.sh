#!/bin/bash
set -e
for v in "-DDSDTM_SA_MINB4=4" "-DDSDTM_SA_MINB4=3"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E "sparse_align_kernelILi4" -A3 | grep -E "Used|spill" | head -2
  for w in 4; do timeout 120 python scripts/prof_step.py --pairs 2368 --steps 3 --direct --wpp $w 2>&1 | tail -2 | head -1 | sed "s/^/wpp=$w /"; done
done
