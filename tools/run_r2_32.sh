#!/bin/bash
mkdir -p gpurun_out
for o in "side_stream=0" "hot_x=0" ""; do
timeout 900 python bench.py --steps 10 --no-configs --no-cpu-baseline --opts "$o" > gpurun_out/p32_bench.json 2> gpurun_out/p32_bench.err
echo "### opts: $o"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/p32_bench.json').read().splitlines() if l.startswith('{')][-1])
print(d['ms_per_step'], 'e2e ms/step', d['e2e']['ms_per_step'], 'serial', d['e2e']['serial_ms_per_step'])
PY
done
