#!/bin/bash
# 2 GPUs: multi-GPU tests with the table kernel in the multicast / peer-store exchange; bench N=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_power_gpu.py tests/test_driver_gpu.py -x -q -m gpu > gpurun_out/p42_pytest.txt 2>&1; tail -3 gpurun_out/p42_pytest.txt
for x in auto p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 50 --warmup 5 --exchange $x --no-configs --no-cpu-baseline > gpurun_out/p42_bench2_$x.json 2> gpurun_out/p42_bench2_$x.err
python - <<P
import json
d=json.loads(open("gpurun_out/p42_bench2_$x.json").read().strip().splitlines()[-1])
print("N=2 $x", d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["kernel"], d["config"].get("exchange"), d["parity"]["ok"], d["parity"]["x_replicas_bit_identical"], "e2e", d["e2e"]["value"])
P
done
timeout 600 ./bin/spmv synthetic:c5 merge --iters 10 --x random --power 50 --gpus 2 > gpurun_out/p42_main2.txt 2>&1; grep -A2 "Time cost\|Power iteration" gpurun_out/p42_main2.txt | cut -c1-200
