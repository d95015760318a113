#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_power_gpu.py -x -q -m gpu -k "native or two_gpu" > gpurun_out/p11_pytest.txt 2>&1
tail -4 gpurun_out/p11_pytest.txt
timeout 600 ./bin/spmv synthetic:c5:24 merge --iters 5 --x random --power 30 --gpus 2 > gpurun_out/p11_main2.txt 2>&1; tail -4 gpurun_out/p11_main2.txt
timeout 900 ./bin/spmv synthetic:c5 merge --iters 3 --x random --power 30 --gpus 2 > gpurun_out/p11_main2_c5.txt 2>&1; tail -4 gpurun_out/p11_main2_c5.txt
