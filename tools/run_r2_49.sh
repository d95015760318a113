#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29771 bench.py --gpus 4 --steps 100 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/p49_bench4.json 2> gpurun_out/p49_bench4.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/p49_bench4.json").read().strip().splitlines()[-1])
c=d["config"]
print("N=4", d["ms_per_step"], d["value"], c["per_rank"]["kernel_ms"], c["exchange"], d["roofline"]["non_kernel_ms"], d["parity"]["ok"], d["parity"]["x_replicas_bit_identical"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_copies_alone_ms_per_step"), d["clocks"])
P
