#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_power_gpu.py -x -q -m gpu > gpurun_out/p8_pytest_power.txt 2>&1
tail -5 gpurun_out/p8_pytest_power.txt
for ex in auto mc; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --exchange $ex > gpurun_out/p8_bench2_$ex.json 2> gpurun_out/p8_bench2_$ex.err
tail -c 400 gpurun_out/p8_bench2_$ex.err
done
python tools/bench_digest.py gpurun_out/p8_bench2_auto.json gpurun_out/p8_bench2_mc.json
