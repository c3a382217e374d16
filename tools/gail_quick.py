#!/usr/bin/env python
"""Production step time of the GAIL Walker program: best of 5 launches of 200 steps (needs a GPU)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

for name in sys.argv[1:] or ["gail_walker"]:
    tr, buf, irl = bench.build_ours(bench.WORKLOADS[name], seed=1, steps_per_launch=200)
    tr.eval_statistics = {}
    if irl is not None:
        irl.disc_eval_statistics = {}
    best = 1e9
    for i in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bench.run_steps(tr, buf, irl, 200)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            best = min(best, e0.elapsed_time(e1) * 1000 / 200)
    print("%s: %.2f us/step (best of 5 x 200-step launches)" % (name, best), flush=True)
