#!/bin/bash
mkdir -p gpurun_out
( echo "### no fill + lib poke"; timeout 600 python tools/step_kernels.py --steps 10 --opts "hot_x_fill=3" --poke lib 2>&1 | grep -E "rank|_kernel|emset|emcpy|elementwise|vectorized" ) > gpurun_out/p21_poke.txt 2>&1
cat gpurun_out/p21_poke.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
