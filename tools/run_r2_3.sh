#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./bin/l2_gather_probe tex > gpurun_out/p3_probe_tex.txt 2>&1
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_abi.py -x -q -m gpu -k "stream or alias or zero or kinds or golden or edge" > gpurun_out/p3_pytest_stream.txt 2>&1
tail -5 gpurun_out/p3_pytest_stream.txt
timeout 600 python tools/quick_bench.py --configs c1,c2 --kinds stream,vector,merge,auto,cusparse --iters 30 > gpurun_out/p3_quick_c1c2.txt 2>&1
timeout 300 python tools/quick_bench.py --configs c1 --kinds stream,vector --iters 30 --no-flush > gpurun_out/p3_quick_c1_warm.txt 2>&1
for n in 1 3 4; do timeout 300 python tools/quick_bench.py --configs c1 --kinds stream --iters 30 --opts stream_ctas_per_sm=$n; done > gpurun_out/p3_quick_c1_ctas.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/p3_pytest_all.txt 2>&1
tail -3 gpurun_out/p3_pytest_all.txt
