#!/bin/bash
set -e
for v in "-DDSDTM_SA_CVT=0" "-DDSDTM_SA_CVT=1" "-DDSDTM_SA_CVT=2"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/sparse_align.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E "sparse_align_kernelILi4" -A3 | grep -E "Used|spill" | head -2
  cuobjdump -sass dsdtm_b200/lib/sparse_align.o | grep -c "I2F" || true
  timeout 120 python scripts/prof_step.py --pairs 2072 --steps 3 --direct --wpp 4 2>&1 | tail -2 | head -1 | cut -c1-130
done
