#!/bin/bash
mkdir -p gpurun_out
for s in 3 4 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --e2e-slots $s > gpurun_out/p54_bench_s$s.json 2> gpurun_out/p54_bench_s$s.err
python - <<P
import json
d=json.loads(open("gpurun_out/p54_bench_s$s.json").read().strip().splitlines()[-1])
print("slots $s: e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "copies alone", d["e2e"]["host_copies_alone_ms_per_step"], "kernel", d["roofline"]["kernel_ms"])
P
done
