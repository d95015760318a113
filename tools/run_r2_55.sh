#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "table" > gpurun_out/p55_pytest.txt 2>&1; tail -1 gpurun_out/p55_pytest.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:merge_tile_table -s 2 -c 1 -f -o gpurun_out/fin_c5_merge_table python tools/prof_one.py --config c5 --kind merge --iters 4 --opts assume_static_pattern=1 > gpurun_out/fin_c5_merge_table_ncu.txt 2>&1
tail -1 gpurun_out/fin_c5_merge_table_ncu.txt
