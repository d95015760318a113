#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "static_pattern or hot_x or table" > gpurun_out/p41_pytest.txt 2>&1; tail -3 gpurun_out/p41_pytest.txt
for t in 0 -1; do for rep in 1 2; do
timeout 600 python bench.py --steps 30 --no-configs --no-cpu-baseline --opts hot_x_table=$t > gpurun_out/p41_bench_t$t.json 2> gpurun_out/p41_bench_t$t.err
python - <<P
import json
d=json.loads(open("gpurun_out/p41_bench_t$t.json").read().strip().splitlines()[-1])
print("hot_x_table=$t", d["ms_per_step"], d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial_ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
P
done; done
