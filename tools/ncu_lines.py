#!/usr/bin/env python
"""Per-source-line stall-sample totals of one ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py report.ncu-rep [top_n]
Reads `ncu --page source --print-source sass,cuda --csv`: SASS rows follow the CUDA line they belong to."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, si = None, None, None
per_line = defaultdict(lambda: [0, defaultdict(int)])
text = {}
cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        text[cur] = r[1].strip()
        continue
    try:
        s = int(r[si])
    except ValueError:
        continue
    if cur is None or s == 0:
        continue
    per_line[cur][0] += s
    for i, c in stall_cols:
        try:
            per_line[cur][1][c] += int(r[i])
        except (ValueError, IndexError):
            pass
tot = sum(v[0] for v in per_line.values())
print("total samples", tot)
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(v[1].items(), key=lambda kv: -kv[1])[:3]
    print("%6d %5.1f%%  %s:%d  [%s]  %s" % (v[0], 100.0 * v[0] / tot, k[0], k[1], " ".join("%s=%d" % (c[6:], n) for c, n in st), text[k][:110]))
