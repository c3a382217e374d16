#!/bin/bash
# A/B of engine builds x environment switches on ONE GPU box (development aid):
#   tools/ab_env.sh "<workloads>" "<lib1.so lib2.so ...>" "<ENV=VAL,ENV=VAL ...>"   -> one phase profile per (lib, env) pair
W="$1"; LIBS="$2"; ENVS="$3"
for lib in $LIBS; do
  for e in $ENVS; do
    echo "######## $lib  [$e]"
    env $(echo $e | tr ',' ' ') ILSW_LIB="$(realpath $lib)" python tools/phase_profile.py $W --no-tile-stamps 2>&1 | cut -c1-300
  done
done
