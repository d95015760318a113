#!/bin/bash
mkdir -p gpurun_out
for v in "" t6x2 t8x2 t4x3; do
  echo "### variant ${v:-default(8x1)}"
  if [ -n "$v" ]; then export SPMVB200_LIB=$PWD/tools/variants/$v.so; else unset SPMVB200_LIB; fi
  timeout 900 python tools/table_sweep.py --configs c3,c5 --sizes 0,99,131,163,195,226 --iters 8
done > gpurun_out/p39_sweep.txt 2>&1; cat gpurun_out/p39_sweep.txt
