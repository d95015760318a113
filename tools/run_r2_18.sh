#!/bin/bash
mkdir -p gpurun_out
for o in "" "hot_x_pdl=0"; do
  echo "### opts: $o"
  timeout 600 python tools/step_kernels.py --steps 10 --opts "$o" 2>&1 | grep -E "rank|_kernel|emset|emcpy"
done > gpurun_out/p18_pdl.txt 2>&1
cat gpurun_out/p18_pdl.txt
timeout 600 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "hot_x" 2>&1 | tail -2
