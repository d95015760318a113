#!/bin/bash
# 1 GPU: full GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p40_pytest.txt 2>&1; tail -4 gpurun_out/p40_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p40_smoke.txt 2>&1; tail -2 gpurun_out/p40_smoke.txt
timeout 1200 python bench.py > gpurun_out/p40_bench.json 2> gpurun_out/p40_bench.err; tail -c 600 gpurun_out/p40_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/p40_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["kernel"], d["parity"]["ok"], d["clocks"])
for c,v in d["configs"].items():
    if c.startswith("_"): continue
    print(c, v["us"], v["us_static_pattern"], v["static_pattern_plan"], v["cusparse_us"], round(v["frac"],3), round(v["frac_static_pattern"],3), v["parity"]["ok"])
P
