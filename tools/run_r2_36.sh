#!/bin/bash
# 8 GPUs: native power iteration (main.cu --power) at 4 and 8 GPUs with peer stores / multicast;
# bench.py at N=2 and N=4 with the multicast exchange (p2p numbers are in p34 / earlier runs)
mkdir -p gpurun_out
python -m pytest tests/test_power_gpu.py tests/test_driver_gpu.py -x -q -m gpu -k "native or power" > gpurun_out/p36_pytest.txt 2>&1; tail -3 gpurun_out/p36_pytest.txt
for n in 8 4; do for x in 0 1; do
  SPMVB200_OPTS=power_exchange=$x timeout 600 ./bin/spmv synthetic:c5 merge --iters 2 --x random --power 50 --gpus $n > gpurun_out/p36_main${n}_x$x.txt 2>&1
  grep -A2 "Power iteration" gpurun_out/p36_main${n}_x$x.txt | cut -c1-200
done; done
for n in 2 4; do for x in mc p2p; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29741 \
    bench.py --gpus $n --steps 50 --warmup 5 --exchange $x --no-configs --no-cpu-baseline --e2e-steps 6 > gpurun_out/p36_bench${n}_$x.json 2> gpurun_out/p36_bench${n}_$x.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/p36_bench${n}_$x.json").read().strip().splitlines()[-1])
    print("bench N=$n exchange $x:", d["ms_per_step"], "ms/step", d["value"], d["roofline"]["kernel_ms"], d["config"].get("exchange"))
except Exception as e: print("bench N=$n $x failed", e)
P
done; done
