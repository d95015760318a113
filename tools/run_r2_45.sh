#!/bin/bash
# 8 GPUs, lean: the bench line at N=8 with the table kernel in the multicast exchange; native driver at 8
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 8 --steps 100 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/p45_bench8.json 2> gpurun_out/p45_bench8.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/p45_bench8.json").read().strip().splitlines()[-1])
c=d["config"]
print("N=8", d["ms_per_step"], d["value"], c["per_rank"]["kernel_ms"], c["exchange"], d["roofline"]["kernel"], d["roofline"]["non_kernel_ms"], d["parity"]["ok"], d["parity"]["x_replicas_bit_identical"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_link_floor_ms_per_step"), d["clocks"])
P
timeout 600 ./bin/spmv synthetic:c5 merge --iters 2 --x random --power 50 --gpus 8 > gpurun_out/p45_main8.txt 2>&1; grep -A2 "Power iteration" gpurun_out/p45_main8.txt | cut -c1-200
