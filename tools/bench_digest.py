"""Print the few numbers of a bench.py JSON line that matter when comparing runs.

    python tools/bench_digest.py gpurun_out/a.log gpurun_out/b.log ...
"""
import json
import sys

for path in sys.argv[1:]:
    try:
        line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception as e:  # noqa: BLE001
        print(f"{path}: no JSON line ({e})")
        continue
    cfg = d.get("config", {})
    pr = cfg.get("per_rank", {})
    print(f"{path}: n_gpus {d.get('n_gpus')}  {d.get('ms_per_step'):.4f} ms/step  {d.get('value'):.1f} {d.get('unit')}"
          f"  exchange {cfg.get('exchange')}  e2e {((d.get('e2e') or {}).get('value'))}")
    if pr:
        print("   kernel_ms", pr.get("kernel_ms"))
        print("   rows     ", [int(v) for v in pr.get("rows", [])])
        print("   nnz      ", [int(v) for v in pr.get("nnz", [])])
    for r in cfg.get("rebalance", []) or []:
        print("   rebalance: before", r.get("local_ms_before") or r.get("kernel_ms_before"))
