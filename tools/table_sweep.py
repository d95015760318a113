"""Shared-memory table of the persistent merge tile kernel (merge_tile_table_kernel): time per SpMV
against the kernel's dynamic shared memory (tiles in flight + table), table off as the baseline;
y must not change by a bit.   python tools/table_sweep.py [--configs c3,c5] [--sizes 0,99,131,...]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from spmv_samples_b200 import generate as gen, spmv  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--configs", default="c3,c5")
p.add_argument("--sizes", default="0,99,131,163,195,226")   # KB of dynamic shared memory; 0 = table off
p.add_argument("--iters", type=int, default=10)
p.add_argument("--f64", action="store_true")
p.add_argument("--opts", default="")
a = p.parse_args()
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    spmv.set_option(k, int(v))
spmv.set_option("assume_static_pattern", 1)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for cfg in a.configs.split(","):
    if a.f64:
        gen.CONFIGS[cfg]["dtype"] = torch.float64
    m = gen.make_config(cfg)
    x = gen.gen_x(m.n_cols, 1, m.Ax.dtype)
    y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")
    y_ref = None
    print(f"== {cfg} rows={m.n_rows} nnz={m.nnz} {m.Ax.dtype} offsets {m.Ap.dtype}", flush=True)
    for kb in [int(s) for s in a.sizes.split(",")]:
        spmv.release_cache()
        spmv.set_option("hot_x_table", 1 if kb else 0)
        if kb:
            spmv.set_option("hot_x_table_bytes", kb << 10)
        for _ in range(3):
            spmv.SpMV("merge", m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            spmv.SpMV("merge", m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        hx = spmv.hot_x_info(m.Aj)
        same = ""
        if y_ref is None:
            y_ref = y.clone()
        else:
            same = "  y bit-identical" if torch.equal(y, y_ref) else "  Y DIFFERS"
        print(f"   smem {kb:3d} KB: {ts[len(ts)//2]:9.1f} us (min {ts[0]:9.1f})  hot {hx['hot_columns']} cols "
              f"{100*hx['hot_share']:.1f}%  table {hx['table_columns']} cols {100*hx['table_share']:.1f}%{same}", flush=True)
    spmv.release_cache()
    spmv.set_option("hot_x_table", -1)
    spmv.set_option("hot_x_table_bytes", -1)
    del m, x, y, y_ref
    torch.cuda.empty_cache()
