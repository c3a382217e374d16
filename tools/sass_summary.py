#!/usr/bin/env python
"""SASS evidence for the Blackwell-specific instructions of the shipped library (no GPU needed):
    python tools/sass_summary.py > profiles/r2_sass_summary.txt
Per kernel: counts of the tcgen05 / TMEM / TMA / mbarrier / multicast mnemonics (names as in B200_PROFILING.md) next to the
legacy tensor-core and cp.async ones, plus the first occurrence of each with its address as a spot check."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ilswiss_b200", "csrc", "libilswiss_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([
    ("UTCMMA / UTCHMMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"),
    ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"), ("UTCATOMSWS / TMEM alloc", r"\bUTCATOMSWS"),
    ("UTMALDG (cp.async.bulk.tensor load)", r"\bUTMALDG"), ("UTMASTG (tensor store)", r"\bUTMASTG"),
    ("UBLKCP (cp.async.bulk 1-D)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("ELECT", r"\bELECT\b"),
    ("HMMA.1688.F32.TF32 (mma.sync)", r"\bHMMA\.1688\.F32\.TF32"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
    ("REDG ... SYS (peer / multicast signal)", r"\bREDG\.[A-Z0-9.]*SYS"), ("CCTL.IVALL", r"\bCCTL\.IVALL"),
    ("LDL / STL (local memory)", r"\b(LDL|STL)\b")])
cur, counts, first = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    for name, pat in PAT.items():
        if re.search(pat, line):
            counts[cur][name] += 1
            first.setdefault((cur, name), line.strip()[:110])
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip().split("(")[0]
print("# cuobjdump -sass %s  (sm_100a), mnemonic counts per kernel -- tools/sass_summary.py" % os.path.relpath(lib, ROOT))
for k, c in counts.items():
    print("\n== %s" % demangle(k))
    for name in PAT:
        if c[name]:
            print("  %6d  %-42s first: %s" % (c[name], name, first[(k, name)]))
