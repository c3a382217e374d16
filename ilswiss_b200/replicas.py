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
    handles = gather_handles(trainer.engine.replica_export(), group)
    trainer.engine.replica_connect(rank, world, handles)
    dist.barrier(group)
    return trainer
