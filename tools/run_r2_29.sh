#!/bin/bash
mkdir -p gpurun_out
for v in default o64b128; do
  if [ $v = default ]; then L=spmv_samples_b200/libspmvb200.so; else L=tools/variants/$v.so; fi
  echo "#### variant $v"
  SPMVB200_LIB=$PWD/$L timeout 600 python tools/quick_bench.py --configs c5 --kinds merge --iters 10 --opts hot_x=1 2>&1 | grep -E "merge|hot-x"
  SPMVB200_LIB=$PWD/$L timeout 600 python tools/step_kernels.py --steps 10 2>&1 | grep -E "rank|_kernel"
done > gpurun_out/p29_o64b128.txt 2>&1
cat gpurun_out/p29_o64b128.txt
