#!/bin/bash
mkdir -p gpurun_out
for v in default s_c1536_s3 s_c1536_s4 s_c1536_s6 s_r128_c768_s4 s_r128_c768_s8 s_r512_c3072_s3; do
  if [ $v = default ]; then L=spmv_samples_b200/libspmvb200.so; else L=tools/variants/$v.so; fi
  for n in 2 3 4 6 8; do
    echo "#### variant $v ctas_per_sm $n"
    SPMVB200_LIB=$PWD/$L timeout 300 python tools/quick_bench.py --configs c1 --kinds stream --iters 30 --opts stream_ctas_per_sm=$n 2>&1 | grep -E "stream|FAILED|Error"
  done
done > gpurun_out/p5_stream_sweep.txt 2>&1
