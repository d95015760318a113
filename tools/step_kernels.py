"""Kernel-level timeline of the row-sharded power iteration (fused norm exchange path): every
rank runs a few steps under the torch profiler (CUPTI sees the library's launches) and prints,
per kernel of a step, its mean duration and the mean idle time in front of it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/step_kernels.py [--workload c5] [--steps 10] [--rebalance 2] [--exchange auto] [--ranks 0,7]
"""
import argparse
import collections
import re
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spmv_samples_b200 import generate  # noqa: E402
from spmv_samples_b200.dist import PowerIteration, init_distributed, shard_rows  # noqa: E402


def summarize(evs, steps, head):
    dur = collections.OrderedDict()
    gap = collections.defaultdict(list)
    prev_end = None
    for e in evs:
        m = re.search(r"(\w+_kernel\w*|[Mm]emset|[Mm]emcpy\w*|nccl\w+)", e.name)
        name = (m.group(1) if m else e.name)[:40]
        dur.setdefault(name, []).append(e.time_range.end - e.time_range.start)
        if prev_end is not None:
            gap[name].append(max(0.0, e.time_range.start - prev_end))
        prev_end = e.time_range.end
    span = (evs[-1].time_range.end - evs[0].time_range.start) / steps if evs else 0.0
    lines = [f"{head}: {span:.1f} us per step over {steps} steps"]
    for name, d in dur.items():
        g = gap.get(name, [0.0])
        lines.append(f"   {name:40s} x{len(d) / steps:4.1f}/step  {sum(d) / len(d):9.1f} us each  "
                     f"idle before {sum(g) / max(len(g), 1):7.1f} us")
    return lines


def plain_spmv_timeline(args, m):
    """Back-to-back plain SpMV(kind) calls (no static flag: the partition kernel runs every time)."""
    from spmv_samples_b200 import spmv
    x = generate.gen_x(m.n_cols, 1, m.Ax.dtype)
    y = torch.empty(m.n_rows, dtype=m.Ax.dtype, device="cuda")
    for _ in range(5):
        spmv.SpMV(args.spmv, m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            spmv.SpMV(args.spmv, m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    print("\n".join(summarize(evs, args.steps, f"SpMV({args.spmv}) on {m.name}")), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--override", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rebalance", type=int, default=2)
    ap.add_argument("--exchange", default="auto")
    ap.add_argument("--ranks", default="")
    ap.add_argument("--opts", default="")
    ap.add_argument("--spmv", default="", help="profile plain SpMV(kind) calls on the workload instead of power steps")
    ap.add_argument("--poke", default="",
                    help="experiment: a tiny unrelated kernel right before every step's SpMV: torch | lib")
    args = ap.parse_args()
    rank, world, _ = init_distributed()
    from spmv_samples_b200 import spmv
    for kv in filter(None, args.opts.split(",")):
        k, v = kv.split("=")
        spmv.set_option(k, int(v))
    import torch.distributed as dist
    gm = generate.make_config(args.workload, 1592635904, scale_override=args.override or None)
    if args.spmv:
        plain_spmv_timeline(args, gm)
        return
    it = PowerIteration(shard_rows(gm, rank, world), gm.n_rows, exchange=args.exchange)
    for _ in range(args.rebalance if world > 1 else 0):
        for _ in range(3):
            it.step()
        it.rebalance(gm, steps=5)
    del gm
    for _ in range(8):
        it.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    poke = torch.zeros(1024, device="cuda")
    from spmv_samples_b200 import _lib
    pk_s = torch.ones(1, dtype=torch.float64, device="cuda")
    pk_a = torch.ones(1, dtype=torch.float32, device="cuda")
    tiny = generate.uniform_rows(4096, 4096, 16, 3)
    tx = generate.gen_x(4096, 1)
    ty = torch.empty(4096, device="cuda")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            if args.poke == "torch":
                poke.add_(1.0)
            elif args.poke == "lib":
                generate.uniform_pm1(1024, 1, 2)
            elif args.poke == "power":     # a kernel of csrc/power.cu (a translation unit without CUB)
                _lib.check(_lib.lib().spmvb200_inv_sqrt(32, pk_s.data_ptr(), pk_a.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream), "inv_sqrt")
            elif args.poke == "vector":    # a kernel of csrc/vector.cu
                spmv.SpMV("vector", tiny.n_rows, tiny.n_cols, tiny.nnz, tiny.Ap, tiny.Aj, tiny.Ax, tx, ty)
            it.step()
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    dur = collections.OrderedDict()
    gap = collections.defaultdict(list)
    prev_end = None
    for e in evs:
        m = re.search(r"(\w+_kernel\w*|[Mm]emset|[Mm]emcpy\w*|nccl\w+)", e.name)
        name = (m.group(1) if m else e.name)[:40]
        dur.setdefault(name, []).append(e.time_range.end - e.time_range.start)
        if prev_end is not None:
            gap[name].append(max(0.0, e.time_range.start - prev_end))
        prev_end = e.time_range.end
    span = (evs[-1].time_range.end - evs[0].time_range.start) / args.steps if evs else 0.0
    lines = [f"rank {rank}/{world} rows {it.shard.csr.n_rows} nnz {it.shard.csr.nnz} exchange {it.exchange}: "
             f"{span:.1f} us per step over {args.steps} steps"]
    for name, d in dur.items():
        g = gap.get(name, [0.0])
        lines.append(f"   {name:40s} x{len(d) / args.steps:4.1f}/step  {sum(d) / len(d):9.1f} us each  "
                     f"idle before {sum(g) / max(len(g), 1):7.1f} us")
    want = [int(v) for v in args.ranks.split(",")] if args.ranks else list(range(world))
    for r in range(world):
        if r == rank and r in want:
            print("\n".join(lines), flush=True)
        if world > 1:
            dist.barrier()
    it.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
