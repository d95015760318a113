"""SpMM (k right-hand sides at once) against k separate SpMVs, per configuration."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmv_samples_b200 import generate as gen, spmv
p = argparse.ArgumentParser()
p.add_argument("--configs", default="c1,c2,c3,c4")
p.add_argument("--iters", type=int, default=10)
p.add_argument("--opts", default="", help="library options, name=value,... (e.g. spmm_by_columns=1)")
a = p.parse_args()
for kv in filter(None, a.opts.split(",")):
    name, v = kv.split("=")
    spmv.set_option(name, int(v))
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sorted(ts)[len(ts) // 2]


for cfg in a.configs.split(","):
    m = gen.make_config(cfg)
    x = gen.gen_x(m.n_cols, 1, m.Ax.dtype)
    y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")
    t1 = timeit(lambda: spmv.SpMV("auto", m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y))
    print(f"== {cfg} nnz={m.nnz}: SpMV(auto) {t1*1e6:9.1f} us  {2*m.nnz/t1/1e9:8.1f} GFLOP/s", flush=True)
    for k in (2, 4, 8):
        X = gen.uniform_pm1(m.n_cols * k, 7, 2, m.Ax.dtype).view(m.n_cols, k)
        Y = torch.empty(m.n_rows, k, dtype=m.Ax.dtype, device="cuda")
        tk = timeit(lambda: spmv.spmm(m.Ap, m.Aj, m.Ax, X, Y))
        print(f"   k={k}: SpMM {tk*1e6:9.1f} us  {2*m.nnz*k/tk/1e9:8.1f} GFLOP/s   vs k SpMVs {k*t1*1e6:9.1f} us  -> {k*t1/tk:4.2f}x", flush=True)
        del X, Y
    del m, x, y
    torch.cuda.empty_cache()
