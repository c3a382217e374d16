#!/usr/bin/env python
"""Per-phase timing of the persistent engine kernel (in-kernel %globaltimer stamps of the last
step of a launch).  Usage: python tools/phase_profile.py [workload ...]   (needs a GPU)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["sac_hopper"]
    for a in sys.argv[1:]:
        if a.startswith("--precision="):
            os.environ["ILSW_GEMM_PRECISION"] = a.split("=")[1]
        if a.startswith("--ctas="):
            os.environ["ILSW_CTAS_PER_SM"] = a.split("=")[1]
    print("gemm precision mode:", os.environ.get("ILSW_GEMM_PRECISION", "default"), " ctas/SM:", os.environ.get("ILSW_CTAS_PER_SM", "auto"))
    for name in names:
        w = bench.WORKLOADS[name]
        tr, buf, irl = bench.build_ours(w, seed=1, steps_per_launch=200)
        tr.eval_statistics = {}
        has_prof = hasattr(tr.engine.lib, "ilsw_trainer_set_profiling")
        # production timing first (tile stamps off): best of 5 launches of 200 steps
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
        print("== %s: %.2f us/step production (best of 5 x 200-step launches, tile stamps off)" % (name, best))
        if has_prof:
            tr.engine.lib.ilsw_trainer_set_profiling(tr.engine.h, 0 if "--no-tile-stamps" in sys.argv else 1)
        if irl is not None:
            irl.disc_eval_statistics = {}
        for _ in range(3):
            bench.run_steps(tr, buf, irl, 200)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bench.run_steps(tr, buf, irl, 200)
        e1.record()
        torch.cuda.synchronize()
        us = tr.engine.phase_times_us()
        desc = tr.engine.describe().strip().split("\n")
        print("== %s: %.1f us/step (launch avg), last-step phase sum %.1f us, %d phases" %
              (name, e0.elapsed_time(e1) * 1000 / 200, us.sum(), len(us)))
        import ctypes as C
        tile = np.zeros((96, 8), dtype=np.uint64)
        tr.engine.lib.ilsw_read_tile_ns(tile.ctypes.data_as(C.c_void_p))
        tile = tile.astype(np.int64)
        cta = np.zeros((96, 304), dtype=np.uint64)
        grid = 148
        if hasattr(tr.engine.lib, "ilsw_read_cta_ns"):
            tr.engine.lib.ilsw_read_cta_ns(tr.engine.h, cta.ctypes.data_as(C.c_void_p), None)
        cta = cta.astype(np.int64)
        pn = np.zeros(2 * 97, dtype=np.uint64)
        tr.engine.lib.ilsw_read_phase_ns(tr.engine.h, pn.ctypes.data_as(C.c_void_p), 2 * 97, None)
        pn = pn.astype(np.int64)
        for i, (t, j, d) in enumerate(zip(us, tr.engine.last_job_us, desc)):
            if t == 0:
                continue
            if "ROW" in d and "GEMM" not in d and tile[i, 0] > 0:
                ts = tile[i, :8]
                n = int((ts > 0).sum())
                d += "   [cta0 row job stages: %s us; job start %.2f us after phase start]" % (
                    " ".join("%.2f" % x for x in np.diff(ts[:n]) / 1000.0), (ts[0] - pn[i]) / 1000.0)
            if "GEMM" in d and tile[i, 4] > 0:
                st = np.diff(tile[i, :5]) / 1000.0
                d += "   [last tile of cta0: issue %.2f  land %.2f  mma %.2f  epi %.2f us]" % tuple(st)
            slow = ""
            if cta[i].max() > 0:
                done = (cta[i] - pn[i]) / 1000.0
                done[cta[i] == 0] = 0
                w = int(done.argmax())
                srt = np.sort(done[done > 0])
                slow = " slowest cta %3d %5.2f, median %5.2f |" % (w, done[w], srt[len(srt) // 2])
            print("  %7.2f us (cta0 jobs %6.2f, barrier+wait %6.2f)%s  %s" % (t, j, t - j, slow, d))


if __name__ == "__main__":
    main()
