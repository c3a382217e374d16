#!/usr/bin/env python
"""Multi-GPU check of the fused in-kernel policy-gradient exchange (run under torchrun, N>=2):
  A. identical replicas  -> bit-identical to a single-replica run ((g+g)/2 == g);
  B. distinct replicas   -> policy stays bit-identical on all ranks, critics differ;
  C. one step vs the R-replica emulation of the CPU oracle (sum of per-replica policy grads / R
     before Adam -- SURVEY.md 8e).
Prints one line 'replica_check OK ...' on rank 0, raises otherwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CFG, G, R, DeviceRun, case_data, case_injection  # noqa: E402
from ilswiss_b200 import replicas  # noqa: E402


def all_equal(t):
    world = dist.get_world_size()
    buf = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(buf, t)
    return all(torch.equal(buf[0], b) for b in buf[1:])


class _T:  # minimal adapter: connect_replicas expects .engine and ._arenas
    def __init__(self, run):
        self.engine = run.eng
        self._arenas = run.nets


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    torch.set_num_threads(2)
    prec = int(os.environ.get("ILSW_CHECK_PRECISION", "0"))
    for name in ("sac_hopper", "td3_hopper"):
        case = CFG.CASES[name]
        inj = case_injection(case)
        # ---- A: identical replicas == single replica, bit for bit
        solo = DeviceRun(case, precision=prec)
        Ls = solo.train(case["steps"], inj)
        rep = DeviceRun(case, precision=prec)
        replicas.connect_replicas(_T(rep))
        Lr = rep.train(case["steps"], inj)
        torch.cuda.synchronize()
        assert np.array_equal(Ls[:, :5], Lr[:, :5], equal_nan=True), (name, "A: losses differ", Ls[:, :5] - Lr[:, :5], Ls[:, 3:5], Lr[:, 3:5])
        for k in ("policy", "qf1", "target_qf1"):
            assert np.array_equal(solo.arena(k), rep.arena(k)), (name, "A", k)
        # ---- B: distinct replicas (different index/eps streams per rank)
        case_b = dict(case)
        inj_b = {k: np.roll(v, rank + 1, axis=0) for k, v in inj.items()}      # per-rank batches
        run = DeviceRun(case_b, precision=prec)
        if rank > 0:   # per-rank critics
            run.nets["qf1"].p.mul_(1.0 + 0.01 * rank)
            run.nets["target_qf1"].p.copy_(run.nets["qf1"].p)
        replicas.connect_replicas(_T(run))
        run.train(case["steps"], inj_b)
        torch.cuda.synchronize()
        assert all_equal(run.nets["policy"].p), (name, "B: policies diverged across ranks")
        assert all_equal(run.nets["policy"].m), (name, "B: policy Adam state diverged")
        assert not all_equal(run.nets["qf1"].p), (name, "B: critics unexpectedly identical")
        # ---- C: one step vs the oracle's R-replica emulation
        if name == "sac_hopper":
            nets = G.build_oracle_nets(case)
            if rank > 0:
                for v in nets["qf1"].p.values():
                    v.mul_(1.0 + 0.01 * rank)
            ora = R.SacAlphaOracle(nets["policy"], nets["qf1"], nets["qf2"], case["act_dim"], **case["sac"])
            data, _ = case_data(case)
            obuf = R.ReplayOracle(case["n_fill"], case["obs_dim"], case["act_dim"])
            obuf.load_bulk(data)
            batch = R.np_to_torch_batch(obuf.get_batch_using_indices(inj_b["idx"][0]))
            pol0 = nets["policy"].clone()
            s = ora.train_step(batch, torch.from_numpy(inj_b["eps_next"][0]), torch.from_numpy(inj_b["eps_cur"][0]))
            g_local = torch.from_numpy(np.concatenate([g.ravel() for g in s["grads"]["policy"]])).cuda()
            gs = [torch.empty_like(g_local) for _ in range(world)]
            dist.all_gather(gs, g_local)
            g_avg = replicas.emulate_replica_average([g.cpu().numpy() for g in gs])
            shapes = [tuple(v.shape) for v in pol0.p.values()]
            parts, off = [], 0
            for shp in shapes:
                n = int(np.prod(shp))
                parts.append(torch.from_numpy(g_avg[off:off + n].reshape(shp)))
                off += n
            R.adam_update(pol0, parts, case["sac"]["policy_lr"], case["sac"].get("beta_1", 0.9))
            one = DeviceRun(case, precision=prec)
            if rank > 0:
                one.nets["qf1"].p.mul_(1.0 + 0.01 * rank)
                one.nets["target_qf1"].p.copy_(one.nets["qf1"].p)
            replicas.connect_replicas(_T(one))
            one.train(1, inj_b)
            diff = np.abs(one.arena("policy") - pol0.flat())
            assert (diff > 1e-5).mean() < 5e-4 and diff.max() < 7e-4, (name, "C", float(diff.max()), float((diff > 1e-5).mean()))
        dist.barrier()
    if rank == 0:
        print("replica_check OK world=%d precision=%d (A: identical==solo bitwise, B: policies bitwise equal across ranks, "
              "C: matches oracle R-replica emulation)" % (world, prec))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
