#!/bin/bash
mkdir -p gpurun_out
( echo "### power iteration, default (side-stream refill)"; timeout 600 python tools/step_kernels.py --steps 10 2>&1 | grep -E "rank|_kernel|emset|emcpy"
  for k in merge light; do echo "### plain SpMV($k) on c3"; timeout 600 python tools/step_kernels.py --workload c3 --spmv $k --steps 10 2>&1 | grep -E "SpMV|_kernel|emset|emcpy"; done
  echo "### plain SpMV(merge) on c2"; timeout 600 python tools/step_kernels.py --workload c2 --spmv merge --steps 10 2>&1 | grep -E "SpMV|_kernel|emset|emcpy"
  echo "### plain SpMV(auto) on c1"; timeout 600 python tools/step_kernels.py --workload c1 --spmv auto --steps 10 2>&1 | grep -E "SpMV|_kernel|emset|emcpy" ) > gpurun_out/p27_timeline.txt 2>&1
cat gpurun_out/p27_timeline.txt
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_power_gpu.py tests/test_driver_gpu.py -x -q -m gpu 2>&1 | tail -3
