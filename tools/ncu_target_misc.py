#!/usr/bin/env python
"""Nsight Compute target for the small kernels next to the engine: ilsw_policy_act_kernel (sampler round trip, env_num 4
and 4096 rows), rb_scatter_kernel (append commit) and the gather kernel on Hopper-shaped rows.
  ncu --set full --clock-control none -k regex:'ilsw_policy_act|rb_scatter|rb_gather' -c 12 -o gpurun_out/misc python tools/ncu_target_misc.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ilswiss_b200.sampler import DevicePolicy  # noqa: E402

w = bench.WORKLOADS["sac_hopper"]
tr, buf, irl = bench.build_ours(w, seed=1, steps_per_launch=10)
dp = DevicePolicy(tr, seed=1)
rs = np.random.RandomState(0)
for n in (4, 4, 4096, 4096):
    dp.get_actions(rs.randn(n, w["O"]))
host = rs.randn(1 << 16, buf.ring.host_w).astype(np.float32)
for _ in range(2):
    buf.ring.append_host(host)
    buf.ring.commit()
for i in range(3):
    buf.ring.sample(1 << 20, 7, i)
idx = torch.randint(0, 1_000_000, (1 << 20,), device="cuda", dtype=torch.int32)
for _ in range(2):
    buf.ring.gather(idx)
torch.cuda.synchronize()
print("misc target done")
