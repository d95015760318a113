#!/bin/bash
mkdir -p gpurun_out
for p in power vector; do
( echo "### no fill + poke $p"; timeout 600 python tools/step_kernels.py --steps 10 --opts "hot_x_fill=3" --poke $p 2>&1 | grep -E "rank|_kernel|emset|emcpy|elementwise|vectorized" )
done > gpurun_out/p22_poke.txt 2>&1
cat gpurun_out/p22_poke.txt
