#!/bin/bash
# 2 GPUs: native multicast exchange (csrc/mcast.cu) against peer stores; main.cu with the static-pattern option
mkdir -p gpurun_out
python -m pytest tests/test_power_gpu.py -x -q -m gpu -k "native" > gpurun_out/p35_pytest.txt 2>&1; tail -5 gpurun_out/p35_pytest.txt
for x in 0 1; do
  SPMVB200_OPTS=power_exchange=$x timeout 600 ./bin/spmv synthetic:c5 merge --iters 20 --x random --power 50 --gpus 2 > gpurun_out/p35_main2_x$x.txt 2>&1
  grep -A3 "Time cost\|Power iteration" gpurun_out/p35_main2_x$x.txt | cut -c1-220
done
