#!/usr/bin/env python
"""Condense tools/phase_profile.py output (development aid): python tools/pp_summary.py file [workload]"""
import re, sys
want = sys.argv[2] if len(sys.argv) > 2 else None
cur = None
for l in open(sys.argv[1]):
    if l.startswith('=='):
        cur = l.split()[1].rstrip(':')
        if not want or cur == want: print(l.strip()[:110])
        continue
    if want and cur != want: continue
    m = re.match(r'\s+([\d.]+) us \(cta0 jobs\s+([\d.]+), barrier\+wait\s+([\d.]+)\)(?: slowest cta\s+(\d+)\s+([\d.]+), median\s+([\d.]+) \|)?\s+phase\s+(\d+) jobs=\s*(\d+)(.*)', l)
    if m:
        st = re.search(r'issue ([\d.]+)\s+land ([\d.]+)\s+mma ([\d.]+)\s+epi ([\d.]+)', l)
        print("ph%2s %6s (cta0 %6s) slowest %3s %6s med %6s jobs %4s %-22s %s" % (m.group(7), m.group(1), m.group(2), m.group(4), m.group(5), m.group(6), m.group(8), "/".join(st.groups()) if st else "", re.sub(r'\[last tile.*', '', m.group(9))[:90]))
