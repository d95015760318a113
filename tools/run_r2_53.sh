#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --no-configs --no-cpu-baseline > gpurun_out/p53_bench.json 2> gpurun_out/p53_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/p53_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(d["ms_per_step"], d["value"], r["frac"], r["traffic"], r["kernel"], d["parity"]["ok"], "e2e", d["e2e"]["value"], d["e2e"]["host_copies_alone_ms_per_step"])
P
