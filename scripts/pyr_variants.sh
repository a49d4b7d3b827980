#!/bin/bash
# Builds pyramid.cu variants on the GPU box and times the pyramid stage of one 4096-pair step (strip8 = --pyr 0, strip4 = --pyr 2).
for v in "-DDSDTM_PYR_RPT=8 -DDSDTM_PYR_UNROLL=2" "-DDSDTM_PYR_RPT=8 -DDSDTM_PYR_UNROLL=4" "-DDSDTM_PYR_RPT=16 -DDSDTM_PYR_UNROLL=4" "-DDSDTM_PYR_RPT=4 -DDSDTM_PYR_UNROLL=2" "-DDSDTM_PYR_RPT=16 -DDSDTM_PYR_UNROLL=2" "-DDSDTM_PYR_RPT=32 -DDSDTM_PYR_UNROLL=2"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/pyramid.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E " error |pyrdown_strip" -A3 | grep -E "error|Used"
  for p in 0 2; do
    timeout 120 python scripts/prof_step.py --pairs 4096 --steps 5 --direct --pyr $p 2>&1 | tail -2 | head -1 | cut -c1-60
  done
done
