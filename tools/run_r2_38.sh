#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "hot_x or table_plan" > gpurun_out/p38_pytest.txt 2>&1; tail -3 gpurun_out/p38_pytest.txt
for lim in 0 4096 -1; do echo "### hot_x_table_limit=$lim"; timeout 900 python tools/table_sweep.py --configs c3,c5 --sizes 0,67,99 --opts hot_x_table_limit=$lim; done > gpurun_out/p38_sweep.txt 2>&1; cat gpurun_out/p38_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_tile_table -s 3 -c 1 -o gpurun_out/p38_c3_table -f python tools/table_sweep.py --configs c3 --sizes 99 --iters 2 > gpurun_out/p38_ncu.log 2>&1; tail -2 gpurun_out/p38_ncu.log
