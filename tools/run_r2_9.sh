#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_fullsize_gpu.py tests/test_power_gpu.py -x -q -m gpu > gpurun_out/p9_pytest.txt 2>&1
tail -5 gpurun_out/p9_pytest.txt
for algo in 0 1; do
  echo "### merge_algo=$algo"
  timeout 600 python tools/quick_bench.py --configs c1,c2,c3,c4 --kinds merge --iters 10 --opts merge_algo=$algo 2>&1 | grep -E "==|merge"
  timeout 600 python tools/quick_bench.py --configs c5 --kinds merge --iters 10 --opts merge_algo=$algo,hot_x=1 2>&1 | grep -E "==|merge|hot-x"
  timeout 600 python tools/quick_bench.py --configs c5 --kinds merge --iters 10 --opts merge_algo=$algo,hot_x=0 2>&1 | grep -E "merge"
done > gpurun_out/p9_merge_ab.txt 2>&1
cat gpurun_out/p9_merge_ab.txt
