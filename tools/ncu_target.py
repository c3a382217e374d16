#!/usr/bin/env python
"""Small deterministic target for Nsight Compute: a few launches of the engine kernel.
Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:ilsw_engine_kernel -s 2 -c 1 \
      -o gpurun_out/prof python tools/ncu_target.py [workload] [steps_per_launch] [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sac_hopper"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
if len(sys.argv) > 3:
    os.environ["ILSW_GEMM_PRECISION"] = sys.argv[3]
w = bench.WORKLOADS[name]
tr, buf, irl = bench.build_ours(w, seed=1, steps_per_launch=steps)
tr.eval_statistics = {}
if irl is not None:
    irl.disc_eval_statistics = {}
for _ in range(4):
    bench.run_steps(tr, buf, irl, steps)
torch.cuda.synchronize()
print("done", name, steps, "launches", tr.engine.kernel_launches)
