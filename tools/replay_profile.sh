#!/bin/bash
# Replay-kernel evidence on one B200 box: the roofline table and one `ncu --set full` capture of rb_gather_kernel per row size.
set -u
OUT=gpurun_out/final; mkdir -p $OUT
python tools/replay_bench.py > $OUT/replay_bench.txt 2>&1
for shape in hopper humanoid; do
  ncu --set full --clock-control none --import-source on -k regex:rb_gather_kernel -s 4 -c 1 -f -o $OUT/gather_$shape \
      python tools/replay_bench.py 18 $shape > $OUT/ncu_gather_$shape.log 2>&1
  ncu -i $OUT/gather_$shape.ncu-rep --page raw --csv > $OUT/gather_${shape}_raw.csv 2>/dev/null
done
cat $OUT/replay_bench.txt | cut -c1-250
