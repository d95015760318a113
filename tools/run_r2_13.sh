#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/pcie_aggregate.py 256 > gpurun_out/p13_pcie2.txt 2>&1; grep GPUs gpurun_out/p13_pcie2.txt
timeout 300 python tools/pcie_aggregate.py 256 > gpurun_out/p13_pcie1.txt 2>&1; grep GPUs gpurun_out/p13_pcie1.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/step_kernels.py --steps 10 > gpurun_out/p13_kernels2.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/p13_kernels2.txt | tail -24
