"""One config, one kind, a few launches: the command that gets wrapped in ncu."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmv_samples_b200 import generate as gen, spmv
p = argparse.ArgumentParser()
p.add_argument("--config", default="c3")
p.add_argument("--kind", default="merge")
p.add_argument("--iters", type=int, default=3)
p.add_argument("--opts", default="")
p.add_argument("--override", type=int, default=0)
p.add_argument("--o64", action="store_true")
a = p.parse_args()
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    spmv.set_option(k, int(v))
if a.o64:
    gen.CONFIGS[a.config]['offset'] = torch.int64
m = gen.make_config(a.config, scale_override=a.override or None)
x = gen.gen_x(m.n_cols, 1, m.Ax.dtype)
y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")
for _ in range(a.iters):
    spmv.SpMV(a.kind, m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
torch.cuda.synchronize()
print("ok", a.config, a.kind, float(y.double().abs().sum()))
