#!/bin/bash
# builds the library on the GPU box once per argument (each argument = a string of nvcc -D switches) and times the sparse alignment
# kernel on the 4096-pair batch: bash scripts/sa_flags.sh "-DDSDTM_SA_LAZY_SHF=0" "-DDSDTM_SA_LAZY_SHF=1"
set -e
for v in "$@"; do
  echo "=== $v"
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E "sparse_align_kernelILi3E" -A3 | grep -E "Used|spill" | head -2
  DSDTM_NVCC_FLAGS="$v" timeout 300 python scripts/sa_sweep.py --combos ${COMBOS:-0:3,0:4,0:10} 2>&1 | tail -3 | cut -c1-75
done
python dsdtm_b200/build.py > /dev/null 2>&1
