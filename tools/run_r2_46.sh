#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_power_gpu.py tests/test_driver_gpu.py -x -q -m gpu > gpurun_out/p46_pytest.txt 2>&1; tail -3 gpurun_out/p46_pytest.txt
