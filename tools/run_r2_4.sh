#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py -x -q -m gpu -k "hot_x or stream or alias or zero" > gpurun_out/p4_pytest_new.txt 2>&1
tail -5 gpurun_out/p4_pytest_new.txt
for mb in 16 32 64; do
  echo "### hot_x_max_bytes = $mb MB"
  timeout 600 python tools/quick_bench.py --configs c5 --kinds merge --iters 10 --opts hot_x=1,hot_x_max_bytes=$((mb<<20))
done > gpurun_out/p4_hotx_c5.txt 2>&1
timeout 600 python tools/quick_bench.py --configs c3 --kinds merge --iters 10 --opts hot_x=1 > gpurun_out/p4_hotx_c3.txt 2>&1
timeout 600 python tools/quick_bench.py --configs c3 --kinds merge --iters 10 --override 26 --o64 >> gpurun_out/p4_hotx_c3.txt 2>&1
timeout 600 python tools/quick_bench.py --configs c3 --kinds merge --iters 10 --override 26 --o64 --opts hot_x=1 >> gpurun_out/p4_hotx_c3.txt 2>&1
timeout 300 python tools/prof_one.py --config c1 --kind stream --iters 3 > gpurun_out/p4_c1_stream_plain.txt 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -c 2 -o gpurun_out/p4_stream_c1 python tools/prof_one.py --config c1 --kind stream --iters 3 > gpurun_out/p4_c1_stream_ncu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/p4_pytest_all.txt 2>&1
tail -3 gpurun_out/p4_pytest_all.txt
