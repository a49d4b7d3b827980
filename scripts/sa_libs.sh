#!/bin/bash
# times the sparse-alignment kernel of prebuilt experiment libraries (scripts/sa_build_variants.py) on the 4096-pair batch; equal sha1 = bit-equal poses
# bash scripts/sa_libs.sh base cvt3 ...
for v in "$@"; do
  echo "=== $v"
  DSDTM_GPU_LIB=dsdtm_b200/lib_exp/$v/libdsdtm_gpu.so timeout 300 python scripts/sa_sweep.py --combos ${COMBOS:-0:3} --steps ${STEPS:-8} 2>&1 | tail -${TAIL:-1}
done
