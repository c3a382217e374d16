#!/bin/bash
# A/B kernel variants on ONE GPU box: tools/ab.sh "<workloads>" lib1.so lib2.so ...  (development aid)
W="$1"; shift
for lib in "$@"; do
  echo "######## $lib"
  ILSW_LIB="$(realpath $lib)" python tools/phase_profile.py $W 2>&1 | cut -c1-400
done
