#!/bin/bash
# 1 GPU: the persistent table kernel -- parity, then time against the table size on c3 and c5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "hot_x or table_plan" > gpurun_out/p37_pytest.txt 2>&1; tail -5 gpurun_out/p37_pytest.txt
timeout 900 python tools/table_sweep.py --configs c3,c5 > gpurun_out/p37_sweep.txt 2>&1; cat gpurun_out/p37_sweep.txt
