#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/table_sweep.py --configs c3 --sizes 0,131,147,163,179,195 > gpurun_out/p44_sweep.txt 2>&1
timeout 900 python tools/table_sweep.py --configs c5 --sizes 83,99,115 >> gpurun_out/p44_sweep.txt 2>&1; cat gpurun_out/p44_sweep.txt
