#!/bin/bash
mkdir -p gpurun_out
for v in default a4 ab2; do
  if [ $v = default ]; then L=spmv_samples_b200/libspmvb200.so; else L=tools/variants/$v.so; fi
  for n in 2 3 4; do
    echo "#### variant $v ctas_per_sm $n"
    SPMVB200_LIB=$PWD/$L timeout 300 python tools/quick_bench.py --configs c1 --kinds stream --iters 30 --opts stream_ctas_per_sm=$n 2>&1 | grep -E "stream|FAILED|Error"
  done
done > gpurun_out/p7_stream_align.txt 2>&1
cat gpurun_out/p7_stream_align.txt
timeout 600 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "stream" 2>&1 | tail -3
