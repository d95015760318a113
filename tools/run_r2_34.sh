#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/pcie_aggregate.py 64 > gpurun_out/p34_pcie8.txt 2>&1; grep GPUs gpurun_out/p34_pcie8.txt
timeout 900 python -m pytest tests/test_power_gpu.py tests/test_driver_gpu.py -x -q -m gpu -k "native_power_iteration_multi or two_gpu or power_iteration_from_one" > gpurun_out/p34_pytest.txt 2>&1
tail -3 gpurun_out/p34_pytest.txt
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/p34_bench$n.json 2> gpurun_out/p34_bench$n.err
tail -c 200 gpurun_out/p34_bench$n.err
python tools/bench_digest.py gpurun_out/p34_bench$n.json
done
timeout 600 ./bin/spmv synthetic:c5 merge --iters 2 --x random --power 50 --gpus 8 > gpurun_out/p34_main8.txt 2>&1; tail -3 gpurun_out/p34_main8.txt
