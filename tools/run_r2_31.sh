#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --no-configs --no-cpu-baseline > gpurun_out/p31_bench1.json 2> gpurun_out/p31_bench1.err
tail -c 300 gpurun_out/p31_bench1.err
python tools/bench_digest.py gpurun_out/p31_bench1.json
timeout 600 python -m pytest tests/test_power_gpu.py tests/test_spmv_gpu.py -x -q -m gpu -k "host or pipelined or hot_x or matrix" 2>&1 | tail -2
