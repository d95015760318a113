#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "table or hot_x or static_pattern" > gpurun_out/p48_pytest.txt 2>&1; tail -3 gpurun_out/p48_pytest.txt
