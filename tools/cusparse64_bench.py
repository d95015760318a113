"""cuSPARSE baseline for the configuration our `cusparse` kind cannot take: int64 row offsets with
int32 column indices (c5) are rejected by cusparseCreateCsr, so the columns are widened to int64
(17 GB more for c5) and the product goes through torch's CSR matmul, which calls cusparseSpMV
with CUSPARSE_INDEX_64I for both arrays.  L2 is not flushed (the operands are far larger).

    python tools/cusparse64_bench.py [--config c5] [--override 0] [--iters 5]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spmv_samples_b200 import generate as gen, spmv  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--config", default="c5")
p.add_argument("--override", type=int, default=0)
p.add_argument("--iters", type=int, default=5)
a = p.parse_args()

m = gen.make_config(a.config, scale_override=a.override or None)
x = gen.gen_x(m.n_cols, 1, m.Ax.dtype)
y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


t_ours = timeit(lambda: spmv.SpMV("auto", m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y))
A = torch.sparse_csr_tensor(m.Ap.long(), m.Aj.long(), m.Ax, size=(m.n_rows, m.n_cols))
out = torch.empty_like(y)
t_cs = timeit(lambda: torch.mv(A, x, out=out))
err = (out.double() - y.double()).abs().max().item()
print(f"{a.config}: rows={m.n_rows} nnz={m.nnz}  ours(auto) {t_ours:8.3f} ms   "
      f"cuSPARSE via torch CSR, 64-bit indices {t_cs:8.3f} ms   ratio {t_cs / t_ours:5.2f}x   max|diff| {err:.3e}")
