#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_power_gpu.py -x -q -m gpu > gpurun_out/p16_pytest.txt 2>&1
tail -4 gpurun_out/p16_pytest.txt
for o in "hot_x_fill=1" "hot_x_fill=2"; do
  echo "### opts: $o"
  timeout 600 python tools/step_kernels.py --steps 10 --opts "$o" 2>&1 | grep -E "rank|_kernel|emset|emcpy"
done > gpurun_out/p16_fill.txt 2>&1
cat gpurun_out/p16_fill.txt
