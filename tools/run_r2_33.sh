#!/bin/bash
mkdir -p gpurun_out
for o in "side_stream=1" "side_stream=0"; do
echo "### SPMVB200_OPTS=$o"
SPMVB200_OPTS=$o timeout 600 ./bin/spmv synthetic:c3 merge --iters 300 2>&1 | grep -E "^\[merge"
SPMVB200_OPTS=$o timeout 600 ./bin/spmv synthetic:c2 merge --iters 300 2>&1 | grep -E "^\[merge" | tail -1
SPMVB200_OPTS=$o timeout 600 python bench.py --steps 50 --no-configs --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('bench ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'non-kernel', d['roofline']['non_kernel_ms'])"
done
