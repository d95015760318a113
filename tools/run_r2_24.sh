#!/bin/bash
mkdir -p gpurun_out
for o in "hot_x_fill=8" "hot_x_fill=9"; do
  echo "### opts: $o"
  timeout 600 python tools/step_kernels.py --steps 10 --opts "$o" 2>&1 | grep -E "rank|_kernel|emset|emcpy"
done > gpurun_out/p24_gap.txt 2>&1
cat gpurun_out/p24_gap.txt
