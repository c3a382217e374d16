#!/bin/bash
# compute-sanitizer over the hand-rolled synchronisation of the engine (grid barrier, mbarrier/TMA pipeline of the tcgen05
# tile, bulk-copy gather): memcheck + synccheck on one SAC, one GAIL, one TD3 and one tcgen05 (batch 512) program,
# racecheck (shared-memory hazards) on the small SAC / GAIL programs.  Run through gpurun; logs -> gpurun_out/<tag>_sanitize_*.
TAG=${1:-san}
mkdir -p gpurun_out
run() {  # tool case precision
  local out=gpurun_out/${TAG}_sanitize_$1_$2.txt
  timeout 900 compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py $2 $3 > $out 2>&1
  echo "rc=$?" >> $out
  echo "== $1 $2: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $out | tr '\n' ' ')"
}
run memcheck sac_ragged 3
run memcheck gail_ragged 3
run memcheck td3_ragged 3
run memcheck sac_hopper_b512_fixed_alpha 3
run synccheck sac_ragged 3
run synccheck sac_hopper_b512_fixed_alpha 3
run racecheck sac_ragged 3
run racecheck gail_ragged 3
