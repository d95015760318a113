#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/p30_pytest.txt 2>&1
tail -4 gpurun_out/p30_pytest.txt
timeout 900 python bench.py > gpurun_out/p30_bench1.json 2> gpurun_out/p30_bench1.err
tail -c 300 gpurun_out/p30_bench1.err
python tools/bench_digest.py gpurun_out/p30_bench1.json
timeout 600 ./bin/spmv synthetic:c3 merge auto cusparse --iters 200 > gpurun_out/p30_main_c3.txt 2>&1; tail -5 gpurun_out/p30_main_c3.txt
timeout 600 ./bin/spmv synthetic:c1 stream vector auto cusparse --iters 2000 > gpurun_out/p30_main_c1.txt 2>&1; tail -6 gpurun_out/p30_main_c1.txt
