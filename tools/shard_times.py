"""Per-shard SpMV time on ONE GPU: emulates the ranks of a P-way row split one after another,
to see the load balance of the merge-path split without paying for P GPUs."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmv_samples_b200 import generate as gen, spmv
from spmv_samples_b200.dist import shard_rows
p = argparse.ArgumentParser()
p.add_argument("--config", default="c5")
p.add_argument("--parts", type=int, default=8)
p.add_argument("--override", type=int, default=0)
p.add_argument("--kind", default="merge")
p.add_argument("--opts", default="")
a = p.parse_args()
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    spmv.set_option(k, int(v))
gm = gen.make_config(a.config, scale_override=a.override or None)
x = gen.gen_x(gm.n_cols, 1, gm.Ax.dtype)
ynext = torch.empty(gm.n_rows, dtype=gm.Ax.dtype, device="cuda")
print(f"{a.config}: rows={gm.n_rows} nnz={gm.nnz} parts={a.parts} kind={a.kind}")
for r in range(a.parts):
    sh = shard_rows(gm, r, a.parts)
    m = sh.csr
    y = ynext[sh.row_begin:sh.row_end]
    for _ in range(2):
        spmv.spmv_ex(a.kind, m.Ap, m.Aj, m.Ax, x, y, n_cols=gm.n_cols)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        spmv.spmv_ex(a.kind, m.Ap, m.Aj, m.Ax, x, y, n_cols=gm.n_cols)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    # how much of a call is the tile kernel, and what the norm pass costs on this shard
    import ctypes
    from spmv_samples_b200 import _lib
    ms, cnt = ctypes.c_double(), ctypes.c_int64()
    spmv.set_option("time_main_kernel", 1)
    _lib.lib().spmvb200_main_kernel_time(ctypes.byref(ms), ctypes.byref(cnt))
    for _ in range(5):
        spmv.spmv_ex(a.kind, m.Ap, m.Aj, m.Ax, x, y, n_cols=gm.n_cols)
    torch.cuda.synchronize()
    _lib.lib().spmvb200_main_kernel_time(ctypes.byref(ms), ctypes.byref(cnt))
    spmv.set_option("time_main_kernel", 0)
    main_ms = ms.value / max(cnt.value, 1)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        _lib.lib().spmvb200_sum_squares(32 if y.element_size() == 4 else 64, y.numel(), y.data_ptr(), ss.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    sumsq_ms = e0.elapsed_time(e1) / 5
    lens = (m.Ap[1:] - m.Ap[:-1])
    print(f"  shard {r}: rows={m.n_rows:10d} nnz={m.nnz:11d} items={m.n_rows + m.nnz:11d} empty={int((lens == 0).sum()):10d} "
          f"max_row={int(lens.max()):8d}  {ts[2]:8.3f} ms  {m.nnz / ts[2] / 1e6:7.1f} Gnnz/s   "
          f"tile kernel {main_ms:7.3f} ms, partition+fixup {ts[2] - main_ms:6.3f} ms, sumsq {sumsq_ms:6.3f} ms", flush=True)
    del sh, m
    torch.cuda.empty_cache()
