#!/bin/bash
# Builds pyramid.cu variants on the GPU box and times the pyramid stage of one 4096-pair step (bulk-staged kernel = --pyr 0).
for v in "-DDSDTM_PYR_BULK_RPT=8 -DDSDTM_PYR_BULK_THREADS=160" "-DDSDTM_PYR_BULK_RPT=8 -DDSDTM_PYR_BULK_THREADS=80" "-DDSDTM_PYR_BULK_RPT=8 -DDSDTM_PYR_BULK_THREADS=320" "-DDSDTM_PYR_BULK_RPT=4 -DDSDTM_PYR_BULK_THREADS=160" "-DDSDTM_PYR_BULK_RPT=4 -DDSDTM_PYR_BULK_THREADS=320" "-DDSDTM_PYR_BULK_RPT=16 -DDSDTM_PYR_BULK_THREADS=160" "-DDSDTM_PYR_BULK_RPT=16 -DDSDTM_PYR_BULK_THREADS=80" "-DDSDTM_PYR_BULK_RPT=12 -DDSDTM_PYR_BULK_THREADS=160" "-DDSDTM_PYR_BULK_RPT=6 -DDSDTM_PYR_BULK_THREADS=240"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/pyramid.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E " error |pyrdown_bulk" -A3 | grep -E "error|Used"
  timeout 120 python scripts/prof_step.py --pairs 4096 --steps 5 --direct --pyr 0 2>&1 | tail -2 | head -1 | cut -c1-60
done
