"""Per-step timeline of the row-sharded power iteration on every rank: how long the local part
(partition + tile kernel + fix-up + sum of squares) takes, how long the tail (norm all-reduce =
step barrier, 1/sqrt) takes, and how both vary from step to step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        tools/step_timeline.py [--workload c5] [--steps 30] [--rebalance 2] [--exchange auto]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spmv_samples_b200 import generate  # noqa: E402
from spmv_samples_b200.dist import PowerIteration, init_distributed, shard_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--override", type=int, default=0)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rebalance", type=int, default=2)
    ap.add_argument("--exchange", default="auto")
    args = ap.parse_args()
    rank, world, _ = init_distributed()
    import torch.distributed as dist
    gm = generate.make_config(args.workload, 1592635904, scale_override=args.override or None)
    it = PowerIteration(shard_rows(gm, rank, world), gm.n_rows, exchange=args.exchange)
    for _ in range(args.rebalance if world > 1 else 0):
        for _ in range(3):
            it.step()
        it.rebalance(gm, steps=5)
    for _ in range(5):
        it.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    it._local_events = []
    ends = []
    for _ in range(args.steps):
        it.step()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ends.append(e)
    torch.cuda.synchronize()
    ev = it._local_events
    it._local_events = None
    local = np.array([a.elapsed_time(b) for a, b in ev])
    tail = np.array([ev[k][1].elapsed_time(ends[k]) for k in range(len(ev))])
    gap = np.array([ends[k].elapsed_time(ev[k + 1][0]) for k in range(len(ev) - 1)])
    whole = np.array([ev[k][0].elapsed_time(ev[k + 1][0]) for k in range(len(ev) - 1)])

    def q(a):
        return f"min {a.min():.3f} med {np.median(a):.3f} mean {a.mean():.3f} max {a.max():.3f}"
    msg = (f"rank {rank}/{world} rows {it.shard.csr.n_rows} nnz {it.shard.csr.nnz} exchange {it.exchange}\n"
           f"   local  {q(local)}\n   tail   {q(tail)}\n   gap    {q(gap)}\n   step   {q(whole)}")
    for r in range(world):
        if r == rank:
            print(msg, flush=True)
        if world > 1:
            dist.barrier()
    it.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
