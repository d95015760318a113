#!/bin/bash
# round-2 GPU call 1: L2 probe, gather-mode variants on c3/c5, per-eviction-class L2 counters
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/p1_smi.txt
nproc >> gpurun_out/p1_smi.txt
timeout 600 ./bin/l2_gather_probe all > gpurun_out/p1_probe_all.txt 2>&1
for v in default g3 g4 g5 g6 g7; do
  if [ $v = default ]; then L=spmv_samples_b200/libspmvb200.so; else L=tools/variants/$v.so; fi
  echo "#### variant $v"
  SPMVB200_LIB=$PWD/$L timeout 600 python tools/quick_bench.py --configs c3,c5 --kinds merge --iters 10
done > gpurun_out/p1_variants.txt 2>&1
M=$(cat tools/l2_metrics.txt)
timeout 300 ./bin/l2_gather_probe ncu > gpurun_out/p1_probe_ncu_plain.txt 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/p1_probe_ncu.csv ./bin/l2_gather_probe ncu > gpurun_out/p1_probe_ncu_run.txt 2>&1
timeout 300 python tools/prof_one.py --config c5 --kind merge --iters 2 > gpurun_out/p1_c5_plain.txt 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:merge_tile --csv --log-file gpurun_out/p1_c5_l2.csv python tools/prof_one.py --config c5 --kind merge --iters 2 > gpurun_out/p1_c5_ncu_run.txt 2>&1
SPMVB200_LIB=$PWD/tools/variants/g4.so timeout 300 python tools/prof_one.py --config c5 --kind merge --iters 2 > gpurun_out/p1_c5g4_plain.txt 2>&1 &&
SPMVB200_LIB=$PWD/tools/variants/g4.so timeout 900 ncu --metrics $M --clock-control none -k regex:merge_tile --csv --log-file gpurun_out/p1_c5g4_l2.csv python tools/prof_one.py --config c5 --kind merge --iters 2 > gpurun_out/p1_c5g4_ncu_run.txt 2>&1
echo done
