#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/step_kernels.py --steps 10 > gpurun_out/p14_kernels1.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$\|Warn\|warn" gpurun_out/p14_kernels1.txt | tail -12
timeout 900 python tools/selector_sweep.py --iters 8 --json gpurun_out/p14_selector.json > gpurun_out/p14_selector.txt 2>&1; cat gpurun_out/p14_selector.txt
