"""Replica scaling: one process per GPU, each with its own replay ring, critics,
alpha and discriminator; ONLY the policy gradient is exchanged, and the exchange happens INSIDE
the step kernel over NVLink peer memory (CUDA IPC mapped receive slots + sequence flags), fused
with the policy Adam -- torch.distributed is used once, at setup, as plumbing."""
import numpy as np
import torch
import torch.distributed as dist


def replica_seed(base_seed, rank):
    """Distinct sampling streams per replica (critics/buffers are independent by design)."""
    return int((base_seed * 1000003 + 7919 * (rank + 1)) % (2 ** 31 - 1))


def emulate_replica_average(per_replica_grads):
    """Oracle-side definition of the exchange step (used by tests): sum in rank order / R."""
    acc = np.zeros_like(np.asarray(per_replica_grads[0], dtype=np.float32))
    for g in per_replica_grads:
        acc = acc + np.asarray(g, dtype=np.float32)
    return acc * np.float32(1.0 / len(per_replica_grads))


def gather_handles(handle, group=None):
    """all_gather of the 64-byte IPC handles (works on gloo and nccl)."""
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, bytes(handle), group=group)
    return out


def connect_replicas(trainer, group=None):
    """Makes `trainer` (ilswiss_b200 SoftActorCritic / TD3) one of dist.get_world_size() replicas:
    broadcasts rank 0's policy (parameters, Adam moments) so all replicas start identical, maps
    every rank's receive slots into this process and switches the engine to replica mode."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return trainer
    arena = trainer._arenas["policy"]
    for t in (arena.p, arena.m, arena.v):
        dist.broadcast(t, src=0, group=group)
    if "target_policy" in trainer._arenas:
        dist.broadcast(trainer._arenas["target_policy"].p, src=0, group=group)
    torch.cuda.synchronize()
    if not _connect_symmetric(trainer, rank, world, group):
        handles = gather_handles(trainer.engine.replica_export(), group)
        trainer.engine.replica_connect(rank, world, handles)
    dist.barrier(group)
    return trainer


def _connect_symmetric(trainer, rank, world, group):
    """Preferred mapping of the exchange buffer: torch's symmetric memory (one allocation per rank, mapped into every process,
    with an NVLS multicast address when the NVSwitch fabric offers one) -- the in-kernel gradient push is then ONE multimem
    store per element.  Every rank must take the same path: the outcome is agreed on with an all_reduce(MIN).  Returns False
    (caller falls back to CUDA IPC peer mappings) when symmetric memory is unavailable (gloo groups, old torch, no fabric)
    or ILSW_REPLICA_SYMM=0."""
    import os
    ok, state = 1, None
    if os.environ.get("ILSW_REPLICA_SYMM", "1") == "0" or dist.get_backend(group) != "nccl":
        ok = 0
    if ok:
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = trainer.engine.replica_buffer_bytes()
            buf = symm.empty((nbytes + 3) // 4, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            buf.zero_()
            torch.cuda.synchronize()
            hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            state = (buf, hdl, nbytes)
        except Exception:       # noqa: BLE001 -- any failure here means "use the IPC path"; agreed on below
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        return False
    buf, hdl, nbytes = state
    mc = int(hdl.multicast_ptr) if os.environ.get("ILSW_REPLICA_MULTICAST", "1") != "0" else 0
    mcf = torch.tensor([1 if mc else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(mcf, op=dist.ReduceOp.MIN, group=group)       # multicast on every rank or on none
    if int(mcf.item()) == 0:
        mc = 0
    trainer.engine.replica_connect_symm(rank, world, [int(p) for p in hdl.buffer_ptrs], mc, nbytes, keepalive=(buf, hdl))
    return True
