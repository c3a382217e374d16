#!/bin/bash
# A/B of replay-kernel builds on one box: tools/replay_ab.sh "<shapes>" lib1.so lib2.so ...
SH="$1"; shift
for lib in "$@"; do
  for s in $SH; do
    ILSW_LIB="$(realpath $lib)" python tools/replay_bench.py 18 $s 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$lib', d['kernel'], d['shape'], d.get('mode', ''), round(d['achieved_gbs']), round(d['frac'], 3))
"
  done
done
