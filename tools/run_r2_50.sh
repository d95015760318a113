#!/bin/bash
mkdir -p gpurun_out
{ echo "### fp64 R-MAT scale 24: marker form (default, smem 0) vs the table kernel (flag form)"; timeout 300 python tools/table_sweep.py --configs c3 --f64 --sizes 0,131,163,195
  echo "### fp64, flag form without a table"; timeout 300 python tools/table_sweep.py --configs c3 --f64 --sizes 0 --opts merge_algo=1; } > gpurun_out/p50_sweep_f64.txt 2>&1; cat gpurun_out/p50_sweep_f64.txt
