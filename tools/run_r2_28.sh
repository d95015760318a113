#!/bin/bash
mkdir -p gpurun_out
( for o in "" "side_stream=0"; do
  echo "### plain SpMV(merge) on c3, opts: $o"; timeout 600 python tools/step_kernels.py --workload c3 --spmv merge --steps 10 --opts "$o" 2>&1 | grep -E "SpMV|_kernel|emset|emcpy"
  done
  echo "### quick_bench (L2 flushed), default then side_stream=0"
  timeout 600 python tools/quick_bench.py --configs c1,c2,c3,c4 --kinds merge,auto --iters 10 2>&1 | grep -E "==|merge|auto"
  timeout 600 python tools/quick_bench.py --configs c3 --kinds merge --iters 10 --opts side_stream=0 2>&1 | grep -E "merge"
  timeout 600 python tools/quick_bench.py --configs c1,c2,c3 --kinds merge,auto --iters 10 --no-flush 2>&1 | grep -E "==|merge|auto"
) > gpurun_out/p28_side.txt 2>&1
cat gpurun_out/p28_side.txt
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_spmm_gpu.py -x -q -m gpu 2>&1 | tail -3
