#!/bin/bash
mkdir -p gpurun_out
( echo "### default"; timeout 600 python tools/step_kernels.py --steps 10 2>&1 | grep -E "rank|_kernel|emset|emcpy"
  echo "### no fill + lib poke (gen.cu kernel; -DNDEBUG now)"; timeout 600 python tools/step_kernels.py --steps 10 --opts "hot_x_fill=3" --poke lib 2>&1 | grep -E "rank|_kernel|emset|emcpy" ) > gpurun_out/p23_fix.txt 2>&1
cat gpurun_out/p23_fix.txt
timeout 900 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "hot_x" 2>&1 | tail -2
