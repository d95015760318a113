#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_power_gpu.py -x -q -m gpu -k "native_power_iteration_multi" > gpurun_out/p12_pytest.txt 2>&1
tail -3 gpurun_out/p12_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/p12_bench8.json 2> gpurun_out/p12_bench8.err
tail -c 300 gpurun_out/p12_bench8.err
python tools/bench_digest.py gpurun_out/p12_bench8.json
timeout 600 ./bin/spmv synthetic:c5 merge --iters 2 --x random --power 50 --gpus 8 > gpurun_out/p12_main8.txt 2>&1; tail -3 gpurun_out/p12_main8.txt
