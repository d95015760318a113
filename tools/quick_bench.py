"""Quick per-config, per-kind device timing (CUDA events, L2 flushed between iterations).
Development aid; bench.py is the contract benchmark."""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmv_samples_b200 import generate as gen, spmv

p = argparse.ArgumentParser()
p.add_argument("--configs", default="c1,c2,c3,c4")
p.add_argument("--kinds", default="merge,vector,light,auto,cusparse")
p.add_argument("--iters", type=int, default=20)
p.add_argument("--opts", default="")  # name=value,name=value
p.add_argument("--no-flush", action="store_true")
p.add_argument("--o64", action="store_true")
p.add_argument("--override", type=int, default=0)
a = p.parse_args()
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    spmv.set_option(k, int(v))
peak = 6452.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for cfg in a.configs.split(","):
    if a.o64:
        gen.CONFIGS[cfg]['offset'] = torch.int64
    m = gen.make_config(cfg, scale_override=a.override or None)
    x = gen.gen_x(m.n_cols, 1, m.Ax.dtype)
    y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")
    st = spmv.row_stats(m.Ap, nnz=m.nnz)
    print(f"== {cfg} rows={m.n_rows} nnz={m.nnz} bytes={m.algorithmic_bytes()} mean={st['mean_row_len']:.2f} "
          f"max={st['max_row_len']} std={st['std_row_len']:.1f} auto->{st['chosen_kind']}/w{st['chosen_width']}", flush=True)
    # yardstick: a plain device copy moving the same number of bytes (half read, half written)
    try:
        half = m.algorithmic_bytes() // 8
        src = torch.empty(half, dtype=torch.float32, device="cuda"); dst = torch.empty_like(src)
        ts = []
        for _ in range(a.iters):
            if not a.no_flush:
                flush.zero_()
            else:
                torch.cuda._sleep(100000)   # keep the GPU busy so the timed launch is already queued
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); dst.copy_(src); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort(); t = ts[len(ts) // 2]
        print(f"   {'(copy)':9s} {t*1e6:10.1f} us  {half*8/t/1e9:8.1f} GB/s   torch copy of the same byte count", flush=True)
        del src, dst
    except Exception as e:
        print("   (copy) failed:", e)
    for kind in a.kinds.split(","):
        try:
            for _ in range(3):
                spmv.SpMV(kind, m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.iters):
                if not a.no_flush:
                    flush.zero_()
                else:
                    torch.cuda._sleep(100000)   # L2 stays warm; hides the launch latency as the flush does
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                spmv.SpMV(kind, m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            ts.sort()
            t = ts[len(ts) // 2]
            gbs = m.algorithmic_bytes() / t / 1e9
            print(f"   {kind:9s} {t*1e6:10.1f} us  {gbs:8.1f} GB/s  {gbs/peak*100:5.1f}% of measured {peak:.0f}  "
                  f"{m.flops()/t/1e9:8.1f} GFLOP/s  (min {ts[0]*1e6:.1f} us)", flush=True)
        except Exception as e:
            print(f"   {kind:9s} FAILED: {e}", flush=True)
    hx = spmv.hot_x_info(m.Aj)
    if hx["hot_columns"]:
        print(f"   hot-x plan: {hx['hot_columns']} columns, {100*hx['hot_share']:.1f}% of the gathers, built in {hx['build_ms']:.1f} ms", flush=True)
    spmv.release_cache()
    del m, x, y
    torch.cuda.empty_cache()
