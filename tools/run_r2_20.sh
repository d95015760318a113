#!/bin/bash
mkdir -p gpurun_out
( echo "### no fill + poke"; timeout 600 python tools/step_kernels.py --steps 10 --opts "hot_x_fill=3" --poke 2>&1 | grep -E "rank|_kernel|emset|emcpy|elementwise|vectorized"
  echo "### hot_x=0 + poke"; timeout 600 python tools/step_kernels.py --steps 10 --opts "hot_x=0" --poke 2>&1 | grep -E "rank|_kernel|emset|emcpy|elementwise|vectorized" ) > gpurun_out/p20_poke.txt 2>&1
cat gpurun_out/p20_poke.txt
