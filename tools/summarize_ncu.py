#!/usr/bin/env python
"""Turns the raw ncu outputs of tools/final_profile.sh (gpurun_out/final/) into the committed summaries under profiles/:
  <tag>_ncu_launches_<workload>.txt  per-kernel launch count / total time / share of the step
  <tag>_ncu_engine_<workload>.txt    key metrics of one `ncu --set full` capture of the engine kernel
  ncu_traffic.json                   DRAM bytes per 1000-step launch (bench.py's roofline.traffic)
usage: python tools/summarize_ncu.py [src_dir=gpurun_out/final] [tag=r1b]"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "final")
tag = sys.argv[2] if len(sys.argv) > 2 else "r1b"
prof = os.path.join(ROOT, "profiles")

# ---- launch list
agg = OrderedDict()
with open(os.path.join(src, "ncu_launches.csv")) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"].split("(")[0][:70]
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + float(r["Metric Value"].replace(",", "")))
total = sum(t for _, t in agg.values())
with open(os.path.join(prof, "%s_ncu_launches_sac_hopper.txt" % tag), "w") as f:
    f.write("# ncu launch list of: python bench.py --steps 3000 --warmup 1000 --e2e-steps 20 --no-cpu-baseline --precision 3\n")
    f.write("# (ncu --metrics gpu__time_duration.sum --clock-control none -c 400; cold-cache, serialised: compare SHARES)\n")
    f.write("%-72s %6s %14s %7s\n" % ("kernel", "count", "total_ns", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-72s %6d %14d %6.2f%%\n" % (k, c, t, 100 * t / total))

# ---- full capture
rows = list(csv.reader(open(os.path.join(src, "engine_full_raw.csv"))))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def val(name):
    i = hdr.index(name)
    return float(vals[i].replace(",", "")), units[i]


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


steps = 50
with open(os.path.join(prof, "%s_ncu_engine_sac_hopper.txt" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on, ilsw_engine_kernel<1>, SAC Hopper B256, %d gradient steps in the "
            "launch, 3xTF32 mode (python tools/ncu_target.py sac_hopper 50 3)\n" % steps)
    for w in want:
        if w in hdr:
            v, u = val(w)
            f.write("%-90s %14.6f %s\n" % (w, v, u))
    rd, wr = to_bytes(*val("dram__bytes_read.sum")), to_bytes(*val("dram__bytes_write.sum"))
    f.write("dram traffic per launch (%d steps): %.3f MB  -> %.1f KB per gradient step (algorithmic bytes/step: 6190.3 KB)\n"
            % (steps, (rd + wr) / 1e6, (rd + wr) / steps / 1e3))
json.dump({"sac_hopper": (rd + wr) / steps * 1000.0,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (50-step launch) scaled to the "
                   "1000-step launch bench.py times"}, open(os.path.join(prof, "ncu_traffic.json"), "w"))
print(open(os.path.join(prof, "%s_ncu_launches_sac_hopper.txt" % tag)).read()[:1500])
print(open(os.path.join(prof, "%s_ncu_engine_sac_hopper.txt" % tag)).read())
