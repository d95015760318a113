#!/bin/bash
# round-2 evidence: ncu --set full per default kernel, per-eviction-class L2 counters (x-only hit rate), bench launch list
mkdir -p gpurun_out
M=$(cat tools/l2_metrics.txt)
cap() {  # name, kernel regex, prof_one args...
  name=$1; rx=$2; shift 2
  timeout 600 python tools/prof_one.py "$@" > gpurun_out/ev_${name}_plain.txt 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/ev_${name} python tools/prof_one.py "$@" > gpurun_out/ev_${name}_ncu.txt 2>&1
  timeout 900 ncu --metrics $M --clock-control none -k regex:$rx -s 1 -c 1 --csv --log-file gpurun_out/ev_${name}_l2.csv python tools/prof_one.py "$@" > /dev/null 2>&1
}
cap c5_merge_hot merge_tile_hot --config c5 --kind merge --iters 3 --opts hot_x=1
cap c3_merge merge_tile_reg --config c3 --kind merge --iters 3
cap c1_stream stream_kernel --config c1 --kind stream --iters 3
cap c2_vector vector_kernel --config c2 --kind vector --iters 3
cap c4_vector vector_kernel --config c4 --kind vector --iters 3
cap c3_light light_kernel --config c3 --kind light --iters 3
cap c2_light light_kernel --config c2 --kind light --iters 3
cap c4_merge_marker merge_tile_reg --config c4 --kind merge --iters 3
timeout 600 python bench.py --steps 3 --warmup 3 --no-configs --e2e-steps 0 --no-cpu-baseline > gpurun_out/ev_bench_plain.txt 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-configs --e2e-steps 0 --no-cpu-baseline > gpurun_out/ev_bench_ncu.txt 2>&1
ls -la gpurun_out/ev_* | head -60
