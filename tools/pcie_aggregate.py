"""Aggregate host<->device bandwidth over the GPUs of a box: every rank copies a pinned buffer to its
GPU (then back, then both at once) at the same time as all the others.  Tells whether the e2e leg of
bench.py at N > 1 can scale with N (per-GPU links) or is bounded by what the host side gives in total.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_aggregate.py [MB]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mb * 1024 * 1024 // 4
h_up = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
h_dn = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(what, reps=10):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if what in ("h2d", "both"):
            with torch.cuda.stream(s_up):
                d_a.copy_(h_up, non_blocking=True)
        if what in ("d2h", "both"):
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / reps


for what in ("h2d", "d2h", "both"):
    run(what, 2)
    dt = run(what)
    if rank == 0:
        per_dir = mb / 1024 / dt
        print(f"{world} GPUs, {mb} MB per GPU per direction, {what:4s}: {dt*1e3:8.3f} ms per round  "
              f"{per_dir:7.1f} GB/s per GPU per direction  {per_dir*world:7.1f} GB/s aggregate per direction", flush=True)
if world > 1:
    dist.destroy_process_group()
