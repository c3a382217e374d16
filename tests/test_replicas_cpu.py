"""CPU (gloo, world_size 2) coverage of the host-side logic of the N>1 path: handle exchange
plumbing, per-replica seeds, and the definition of the exchange step used as the multi-GPU
oracle (sum of per-replica policy gradients in rank order / R)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ilswiss_b200 import replicas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    handle = bytes([rank + 1]) * 64                         # stand-in for a cudaIpcMemHandle_t
    got = replicas.gather_handles(handle)
    assert len(got) == world and all(len(h) == 64 for h in got)
    assert [h[0] for h in got] == [r + 1 for r in range(world)]      # rank order preserved
    # the exchange step's arithmetic: every rank forms the same rank-ordered sum
    g = torch.full((1000,), float(rank + 1)) + torch.arange(1000) * 1e-3
    gs = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gs, g)
    avg = replicas.emulate_replica_average([x.numpy() for x in gs])
    ref = gs[0].clone()
    for x in gs[1:]:
        ref = ref + x
    ref = ref * (1.0 / world)
    assert np.array_equal(avg, ref.numpy())
    t = torch.from_numpy(avg.copy())
    chk = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(chk, t)
    assert all(torch.equal(chk[0], c) for c in chk)          # bit-identical on all ranks
    out[rank] = replicas.replica_seed(12345, rank)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and out[0] != out[1]            # distinct sampling streams


def test_emulated_average_is_rank_ordered_fp32():
    rs = np.random.RandomState(0)
    gs = [rs.randn(1000).astype(np.float32) for _ in range(8)]
    avg = replicas.emulate_replica_average(gs)
    acc = np.zeros(1000, np.float32)
    for g in gs:
        acc = acc + g
    assert np.array_equal(avg, acc * np.float32(0.125)) and avg.dtype == np.float32
