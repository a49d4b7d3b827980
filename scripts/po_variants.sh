#!/bin/bash
# Register caps / CTA shapes of the pose-refinement sweep kernel (one warp per frame): 4096 frames x 300 observations.
for v in "" "-DDSDTM_PO_SWEEP_MINB=1" "-DDSDTM_PO_SWEEP_MINB=3" "-DDSDTM_PO_SWEEP_MINB=4" "-DDSDTM_PO_SWEEP_MINB=5" "-DDSDTM_PO_SWEEP_WARPS=2 -DDSDTM_PO_SWEEP_MINB=8" "-DDSDTM_PO_SWEEP_WARPS=8 -DDSDTM_PO_SWEEP_MINB=2"; do
  echo "=== $v"
  touch dsdtm_b200/csrc/pose_opt.cu
  DSDTM_NVCC_FLAGS="$v" python dsdtm_b200/build.py 2>&1 | grep -E " error " -A3
  grep -n "pose_opt_sweep" -A3 dsdtm_b200/lib/ptxas.log | grep -E "Used|spill" | head -2
  timeout 120 python scripts/prof_pose_opt.py 4096 300 5 2>&1 | head -1
done
touch dsdtm_b200/csrc/pose_opt.cu
python dsdtm_b200/build.py > /dev/null 2>&1
