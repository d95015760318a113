"""Selector sweep: a dozen matrices between "regular" and "power-law", every
kind timed (CUDA events, L2 flushed, median), the selector's choice against the best kind.

    python tools/selector_sweep.py [--iters 10] [--rows 4194304] [--json out.json]

The matrices are built with torch on the device (row-length distribution -> Ap by cumsum, columns
uniform or clustered, values U(-1,1)); the R-MAT ones come from the library's generator.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spmv_samples_b200 import generate as gen, spmv

KINDS = ["merge", "vector", "light", "stream"]


def from_lengths(name, lens, n_cols, seed, band=None):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lens = lens.to(torch.int64).clamp_(min=0)
    n_rows = lens.numel()
    Ap = torch.zeros(n_rows + 1, dtype=torch.int64, device="cuda")
    Ap[1:] = torch.cumsum(lens, 0)
    nnz = int(Ap[-1])
    if band is None:
        Aj = torch.randint(0, n_cols, (nnz,), device="cuda", generator=g, dtype=torch.int32)
    else:   # columns within +-band of the row
        rows = torch.repeat_interleave(torch.arange(n_rows, device="cuda"), lens)
        off = torch.randint(-band, band + 1, (nnz,), device="cuda", generator=g)
        Aj = (rows + off).clamp_(0, n_cols - 1).to(torch.int32)
        del rows, off
    Ax = torch.rand(nnz, device="cuda", generator=g) * 2 - 1
    off_t = torch.int32 if nnz < 2 ** 31 - 8192 else torch.int64
    return gen.Csr(n_rows, n_cols, nnz, Ap.to(off_t), Aj, Ax, name)


def lognormal_lengths(n, mean, sigma, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(n, device="cuda", generator=g)
    mu = torch.log(torch.tensor(float(mean))) - sigma * sigma / 2
    return torch.exp(mu + sigma * z).round()


def family(n):
    g = torch.Generator(device="cuda").manual_seed(99)
    yield from_lengths("uniform len 3", torch.full((n,), 3, device="cuda"), n, 1)
    yield from_lengths("uniform len 8, band 64", torch.full((n,), 8, device="cuda"), n, 2, band=64)
    yield from_lengths("uniform len 16", torch.full((n,), 16, device="cuda"), n, 3)
    yield from_lengths("uniform len 64", torch.full((n // 4,), 64, device="cuda"), n // 4, 4)
    yield from_lengths("len 1..31 uniform", torch.randint(1, 32, (n,), device="cuda", generator=g), n, 5)
    yield from_lengths("lognormal mean 16 sigma 0.5", lognormal_lengths(n, 16, 0.5, 6), n, 6)
    yield from_lengths("lognormal mean 16 sigma 1.0", lognormal_lengths(n, 16, 1.0, 7), n, 7)
    yield from_lengths("lognormal mean 16 sigma 1.5", lognormal_lengths(n, 16, 1.5, 8), n, 8)
    yield from_lengths("lognormal mean 16 sigma 2.0", lognormal_lengths(n, 16, 2.0, 9), n, 9)
    half = torch.where(torch.rand(n, device="cuda", generator=g) < 0.5, 0, 32)
    yield from_lengths("half the rows empty, others 32", half, n, 10)
    few = torch.full((n,), 8, device="cuda")
    few[torch.randint(0, n, (64,), device="cuda", generator=g)] = 200000
    yield from_lengths("len 8 + 64 rows of 200000", few, n, 11)
    yield gen.rmat(22, 16, 12)
    m = gen.rmat(21, 32, 13)
    m.name = "rmat s21 ef32"
    yield m


def median_us(fn, flush, iters):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--iters", type=int, default=10)
    p.add_argument("--rows", type=int, default=1 << 22)
    p.add_argument("--json", default="")
    a = p.parse_args()
    names = {0: "merge", 1: "vector", 2: "light", 5: "stream"}
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    out = []
    print(f"{'matrix':38s} {'mean':>6s} {'max':>8s} {'std':>8s} | " + " ".join(f"{k:>9s}" for k in KINDS)
          + f" | {'auto':>9s} {'chosen':>7s} {'auto/best':>9s}")
    for m in family(a.rows):
        x = gen.gen_x(m.n_cols, 1, m.Ax.dtype)
        y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")
        st = spmv.row_stats(m.Ap, nnz=m.nnz)
        t = {}
        for k in KINDS + ["auto"]:
            if k == "stream" and (st["mean_row_len"] > 64 or st["max_row_len"] > 4096):
                t[k] = float("inf")     # a thread per row: minutes on a hub row
                continue
            call = lambda: spmv.SpMV(k, m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
            for _ in range(2):
                call()
            t[k] = median_us(call, flush, a.iters)
        best = min(KINDS, key=lambda k: t[k])
        ratio = t[best] / t["auto"]
        print(f"{m.name:38s} {st['mean_row_len']:6.1f} {st['max_row_len']:8d} {st['std_row_len']:8.1f} | "
              + " ".join(f"{t[k]:9.1f}" for k in KINDS)
              + f" | {t['auto']:9.1f} {names.get(st['chosen_kind'], '?'):>7s} {ratio:9.3f}", flush=True)
        out.append({"matrix": m.name, "rows": m.n_rows, "nnz": m.nnz, "stats": st, "us": t, "best": best,
                    "chosen": names.get(st["chosen_kind"]), "best_over_auto": ratio})
        spmv.release_cache()
        del m, x, y
        torch.cuda.empty_cache()
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1, default=str)


if __name__ == "__main__":
    main()
