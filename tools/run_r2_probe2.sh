#!/bin/bash
mkdir -p gpurun_out
timeout 600 ./bin/l2_gather_probe v2 > gpurun_out/p2_probe_v2.txt 2>&1
M=$(cat tools/l2_metrics2.txt)
timeout 300 ./bin/l2_gather_probe ncu2 > gpurun_out/p2_probe_ncu2_plain.txt 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:probe_kernel --csv --log-file gpurun_out/p2_probe_ncu2.csv ./bin/l2_gather_probe ncu2 > gpurun_out/p2_probe_ncu2_run.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/p2_pytest.txt 2>&1
tail -3 gpurun_out/p2_pytest.txt
