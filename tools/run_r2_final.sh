#!/bin/bash
# final single-GPU evidence of round 2: GPU suite, smoke, both bench arms, launch list, ncu captures of the table kernel
mkdir -p gpurun_out
M=$(cat tools/l2_metrics.txt)
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/fin_pytest.txt 2>&1; tail -3 gpurun_out/fin_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.txt 2>&1; tail -1 gpurun_out/fin_smoke.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err; tail -c 400 gpurun_out/fin_ref.json; echo
timeout 1500 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; tail -c 300 gpurun_out/fin_bench.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-configs --e2e-steps 0 --no-cpu-baseline > gpurun_out/fin_bench_plain.txt 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --e2e-steps 0 --no-cpu-baseline > gpurun_out/fin_bench_ncu.txt 2>&1
cap() {  # name, kernel regex, prof_one args...
  name=$1; rx=$2; shift 2
  timeout 600 python tools/prof_one.py "$@" > gpurun_out/fin_${name}_plain.txt 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/fin_${name} python tools/prof_one.py "$@" > gpurun_out/fin_${name}_ncu.txt 2>&1
  timeout 900 ncu --metrics $M --clock-control none -k regex:$rx -s 2 -c 1 --csv --log-file gpurun_out/fin_${name}_l2.csv python tools/prof_one.py "$@" > /dev/null 2>&1
}
cap c5_merge_table merge_tile_table --config c5 --kind merge --iters 5 --opts assume_static_pattern=1
cap c3_merge_table merge_tile_table --config c3 --kind merge --iters 5 --opts assume_static_pattern=1
ls -la gpurun_out/fin_* | cut -c30-120
