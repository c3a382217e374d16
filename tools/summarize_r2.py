#!/usr/bin/env python
"""Runs ON the GPU box right after tools/evidence_r2.sh's captures: turns the .ncu-rep files (16 MB each -- too large to
bring back together) into the small text summaries committed under profiles/.
  ncu_engine_<workload>.txt   key metrics of one `ncu --set full` capture of the engine kernel + its hottest source lines
  ncu_misc.txt                one line per captured launch of the small kernels (policy_act, scatter, gather)
  ncu_traffic.json            DRAM bytes per gradient step per workload (bench.py's roofline.traffic)
usage: python tools/summarize_r2.py gpurun_out/r2"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2")
STEPS = {"td3_humanoid": 4, "gail_walker": 10, "sac_hopper": 10, "sac_ant": 10}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))


traffic = {}
for wl, steps in STEPS.items():
    rep = os.path.join(src, "ncu_%s.ncu-rep" % wl)
    if not os.path.exists(rep):
        continue
    rows = raw(rep)
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    with open(os.path.join(src, "ncu_engine_%s.txt" % wl), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:ilsw_engine_kernel -s 2 -c 1: python tools/ncu_target.py %s %d 3\n"
                "# kernel: %s ; %d gradient steps in the captured launch, 3xTF32 mode\n" % (wl, steps, kname, steps))
        get = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                get[w] = (float(vals[i].replace(",", "")), units[i])
                f.write("%-90s %16.6f %s\n" % (w, get[w][0], units[i]))
        if "dram__bytes_read.sum" in get:
            rd = get["dram__bytes_read.sum"][0] * UNIT.get(get["dram__bytes_read.sum"][1], 1)
            wr = get["dram__bytes_write.sum"][0] * UNIT.get(get["dram__bytes_write.sum"][1], 1)
            traffic[wl + "_per_step"] = (rd + wr) / steps
            f.write("dram traffic: %.3f MB per launch -> %.1f KB per gradient step\n" % ((rd + wr) / 1e6, (rd + wr) / steps / 1e3))
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "30"], capture_output=True, text=True).stdout
        f.write("\n# hottest source lines by warp-stall samples (tools/ncu_lines.py; stalls: top three reasons per line)\n" + lines)
    if wl != "sac_hopper":
        os.remove(rep)          # keep one report for source-level reading back home; the others stay as summaries
json.dump(dict(traffic, note="dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the engine kernel, per gradient step"),
          open(os.path.join(src, "ncu_traffic.json"), "w"), indent=1)

rep = os.path.join(src, "ncu_misc.ncu-rep")
if os.path.exists(rep):
    rows = raw(rep)
    hdr, units = rows[0], rows[1]
    cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
    with open(os.path.join(src, "ncu_misc.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none -k regex:'ilsw_policy_act|rb_scatter|rb_gather' : python tools/ncu_target_misc.py\n")
        f.write("# policy_act: 4-row and 4096-row calls; scatter: 65536 Hopper rows; gather: 1M Hopper rows (Philox sample, then caller indices)\n")
        f.write(" | ".join(c + ("" if c == "Kernel Name" else " [%s]" % units[hdr.index(c)]) for c in cols if c in hdr) + "\n")
        for r in rows[2:]:
            f.write(" | ".join(r[hdr.index(c)][:60] for c in cols if c in hdr) + "\n")
    os.remove(rep)
print("summaries:", sorted(os.listdir(src)))
