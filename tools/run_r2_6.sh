#!/bin/bash
mkdir -p gpurun_out
for v in default ab1 ab2; do
  if [ $v = default ]; then L=spmv_samples_b200/libspmvb200.so; else L=tools/variants/$v.so; fi
  for n in 2 4; do
    echo "#### variant $v ctas_per_sm $n"
    SPMVB200_LIB=$PWD/$L timeout 300 python tools/quick_bench.py --configs c1 --kinds stream --iters 30 --opts stream_ctas_per_sm=$n 2>&1 | grep -E "stream|FAILED|Error|copy"
  done
done > gpurun_out/p6_stream_ablate.txt 2>&1
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/p6_bench1.json 2> gpurun_out/p6_bench1.err
tail -c 600 gpurun_out/p6_bench1.err
python tools/bench_digest.py gpurun_out/p6_bench1.json 2>&1 | tail -20
