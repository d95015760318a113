#!/bin/bash
mkdir -p gpurun_out
for o in "" "merge_carveout=-1" "hot_x=0" "hot_x=0,merge_carveout=-1"; do
  echo "### opts: $o"
  timeout 600 python tools/step_kernels.py --steps 10 --opts "$o" 2>&1 | grep -E "rank|_kernel|emset|emcpy"
done > gpurun_out/p15_gap.txt 2>&1
cat gpurun_out/p15_gap.txt
