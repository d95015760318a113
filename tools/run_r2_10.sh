#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_power_gpu.py -x -q -m gpu > gpurun_out/p10_pytest.txt 2>&1
tail -4 gpurun_out/p10_pytest.txt
for algo in 0 1; do
  echo "### merge_algo=$algo"
  timeout 600 python tools/quick_bench.py --configs c2,c3,c4 --kinds merge --iters 10 --opts merge_algo=$algo 2>&1 | grep -E "merge"
  timeout 600 python tools/quick_bench.py --configs c5 --kinds merge --iters 10 --opts merge_algo=$algo,hot_x=1 2>&1 | grep -E "merge|hot-x"
  timeout 600 python tools/quick_bench.py --configs c5 --kinds merge --iters 10 --opts merge_algo=$algo,hot_x=0 2>&1 | grep -E "merge"
done > gpurun_out/p10_merge_ab.txt 2>&1
cat gpurun_out/p10_merge_ab.txt
timeout 600 ./bin/spmv synthetic:c3:20 merge auto stream cusp light_vec cub_merge --iters 20 --x random --power 10 > gpurun_out/p10_main.txt 2>&1; tail -12 gpurun_out/p10_main.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/p10_ref.json 2> gpurun_out/p10_ref.err; tail -c 300 gpurun_out/p10_ref.err; head -c 1500 gpurun_out/p10_ref.json
