#!/usr/bin/env python
"""Probe (torchrun, N >= 2): does torch's symmetric memory give an NVLS multicast mapping on this box?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = symm.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", lr))
h = symm.rendezvous(t, dist.group.WORLD)
print("rank", rank, "multicast", h.has_multicast_support(torch.device("cuda", lr).type, lr) if False else None, hex(h.multicast_ptr), [hex(p) for p in h.buffer_ptrs], "size", h.buffer_size, flush=True)
dist.barrier()
dist.destroy_process_group()
