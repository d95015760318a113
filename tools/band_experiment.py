"""What would column-band blocking buy on c5?  Split the matrix into B column bands (torch
ops, one-off), time the merge SpMV of every band (each gathers only from its slice of x), and
compare the sum with the unbanded SpMV.  Experiment only -- no accumulate pass is timed."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmv_samples_b200 import generate as gen, spmv

p = argparse.ArgumentParser()
p.add_argument("--config", default="c5")
p.add_argument("--override", type=int, default=0)
p.add_argument("--bands", default="2,4,8")
a = p.parse_args()
m = gen.make_config(a.config, scale_override=a.override or None)
x = gen.gen_x(m.n_cols, 1, m.Ax.dtype)
y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")


def timeit(Ap, Aj, Ax, iters=3):
    for _ in range(2):
        spmv.spmv_ex("merge", Ap, Aj, Ax, x, y, n_cols=m.n_cols)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        spmv.spmv_ex("merge", Ap, Aj, Ax, x, y, n_cols=m.n_cols)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


base = timeit(m.Ap, m.Aj, m.Ax)
print(f"{a.config}: rows={m.n_rows} nnz={m.nnz}  unbanded merge {base:.3f} ms", flush=True)
bits = (m.n_cols - 1).bit_length()
for B in [int(b) for b in a.bands.split(",")]:
    shift = bits - (B.bit_length() - 1)
    total = 0.0
    parts = []
    band = (m.Aj >> shift).to(torch.int8)
    for b in range(B):
        mask = band == b
        csum = torch.cumsum(mask, 0, dtype=torch.int64)
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), csum])
        Ap_b = csum[m.Ap.long()].contiguous()
        del csum
        Aj_b = m.Aj[mask].contiguous()
        Ax_b = m.Ax[mask].contiguous()
        del mask
        t = timeit(Ap_b, Aj_b, Ax_b)
        parts.append((int(Aj_b.numel()), t))
        total += t
        del Ap_b, Aj_b, Ax_b
        torch.cuda.empty_cache()
    print(f"  {B} bands: sum {total:.3f} ms ({base / total:.2f}x)  per band (nnz, ms): "
          + ", ".join(f"({n/1e6:.0f}M, {t:.2f})" for n, t in parts), flush=True)
    del band
