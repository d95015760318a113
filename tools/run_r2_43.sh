#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_power_gpu.py -x -q -m gpu -k "static_pattern or hot_x or table or rebalanced or power" > gpurun_out/p43_pytest.txt 2>&1; tail -3 gpurun_out/p43_pytest.txt
timeout 900 python tools/table_sweep.py --configs c3,c5 --sizes 0,67,99,131 > gpurun_out/p43_sweep.txt 2>&1; cat gpurun_out/p43_sweep.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/p43_bench2.json 2> gpurun_out/p43_bench2.err
python - <<P
import json
d=json.loads(open("gpurun_out/p43_bench2.json").read().strip().splitlines()[-1])
c=d["config"]
print("N=2", d["ms_per_step"], d["value"], c["per_rank"], [r["local_ms_before"] for r in c["rebalance"]], c["hot_x"]["table_columns"], c["hot_x"]["table_share"], d["parity"]["ok"], "e2e", d["e2e"]["value"])
P
