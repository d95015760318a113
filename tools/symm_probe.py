"""Probe: does torch symmetric memory give peer pointers and an NVLS multicast pointer here?"""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{lr}")
try:
    h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
except TypeError:
    h = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(getattr(h, "multicast_ptr", 0) or 0), flush=True)
dist.barrier()
dist.destroy_process_group()
