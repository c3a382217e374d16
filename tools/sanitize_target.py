#!/usr/bin/env python
"""Small deterministic targets for compute-sanitizer (tools/sanitize.sh): a couple of engine launches of a parity case
through the C ABI, plus one replay-ring append / commit / gather round.  Usage: sanitize_target.py <case> [precision]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CFG, DeviceRun, case_injection  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sac_ragged"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = dict(CFG.CASES[name], steps=2)
run = DeviceRun(case, precision=prec)
L = run.train(2, case_injection(case))
assert np.isfinite(L[:, :2]).all(), L
run.train_philox(2, seed=3)
idx = torch.arange(0, 64, dtype=torch.int32, device="cuda") % case["n_fill"]
hot, cold = run.ring.gather(idx)
rows = np.zeros((5, run.ring.host_w), np.float32)
run.ring.append_host(rows)
run.ring.commit()
run.ring.sample(96, 7, 1)
torch.cuda.synchronize()
print("sanitize target ok:", name, "precision", prec, "tcgen05" if run.eng.uses_tc5() else "mma.sync", "losses", L[:, :2].tolist())
