#!/usr/bin/env python
"""Replay-ring kernels against the HBM roofline: uniform sample + minibatch gather and the append
scatter, at the four BASELINE.json shapes, on large batches (the per-step batches of 256-1024 rows are launch-latency
bound; this measures what the kernels sustain).  Prints one JSON line per (shape, kernel).

    python tools/replay_bench.py [rows_log2=18] [shape]
Algorithmic bytes: gather = B * stride * 4 read (random rows) + B * stride * 4 written (+ 4 B index bytes);
scatter = n * host_w * 4 read + n * (stride + 4) * 4 written."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilswiss_b200 import engine  # noqa: E402

SHAPES = {"hopper": (11, 3, 1_000_000), "walker": (17, 6, 1_000_000), "ant": (111, 8, 1_000_000), "humanoid": (376, 17, 2_000_000)}


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    B = 1 << lg
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    only = sys.argv[2] if len(sys.argv) > 2 else None
    for name, (O, A, N) in SHAPES.items():
        if only and name != only:
            continue
        ring = engine.ReplayRing(N, O, A)
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        chunk = 250_000
        for s in range(0, N, chunk):
            ring.load_device(torch.randn((min(chunk, N - s), ring.stride), generator=g, device="cuda"))
        B = (1 << lg) * (4 if ring.stride <= 64 else 1)      # small rows: more of them, so the launch ramp does not dominate
        idx = torch.randint(0, N, (B,), generator=g, device="cuda", dtype=torch.int32)
        for label, fn in (("gather(idx)", lambda: ring.gather(idx)), ("sample(philox)", lambda: ring.sample(B, 7, 1))):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.zero_()                                 # L2 flush (untimed)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                del out
            ms = float(np.median(ts))
            by = B * ring.stride * 4 * 2 + B * 4
            print(json.dumps({"kernel": "rb_gather_kernel", "mode": label, "shape": name, "rows": B, "row_bytes": ring.stride * 4,
                              "ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6, "peak_gbs": peak,
                              "frac": by / ms / 1e6 / peak, "note": "includes the output allocation of the Python wrapper (cached allocator)"}))
        n = min(1 << 17, (64 << 20) // (ring.host_w * 4))
        host = np.random.RandomState(0).randn(n, ring.host_w).astype(np.float32)
        ts = []
        for _ in range(5):
            ring.append_host(host)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ring.commit(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        by = n * ring.host_w * 4 + n * (ring.stride + 4) * 4
        print(json.dumps({"kernel": "rb_scatter_kernel", "shape": name, "rows": n, "ms": ms, "algorithmic_bytes": by,
                          "achieved_gbs": by / ms / 1e6, "peak_gbs": peak, "frac": by / ms / 1e6 / peak}))
        del ring
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
