#!/bin/bash
# Round-2 evidence on ONE B200 box (run through gpurun): replay-kernel roofline, per-phase profiles, ncu captures of the
# TD3 (tcgen05) / GAIL / SAC programs and the small kernels, compute-sanitizer over the hand-rolled synchronisation.
# Outputs: gpurun_out/r2/ (summarised into profiles/ by tools/summarize_r2.py).   usage: tools/evidence_r2.sh [steps...]
OUT=gpurun_out/r2; mkdir -p $OUT
STEPS=${@:-replay pp ncu misc launches sanitize summarize}
for s in $STEPS; do case $s in
  replay) timeout 300 python tools/replay_bench.py 18 > $OUT/replay_bench.txt 2>&1
          ILSW_GATHER_LEGACY=1 timeout 300 python tools/replay_bench.py 18 > $OUT/replay_bench_legacy.txt 2>&1 ;;
  pp) timeout 600 python tools/phase_profile.py sac_hopper gail_walker td3_humanoid sac_ant her_td3_pick > $OUT/phase_profile.txt 2>&1 ;;
  ncu) for wl in "td3_humanoid 4" "gail_walker 10" "sac_ant 10" "sac_hopper 10"; do set -- $wl
         timeout 600 ncu --set full --clock-control none --import-source on -k regex:ilsw_engine_kernel -s 2 -c 1 -f -o $OUT/ncu_$1 python tools/ncu_target.py $1 $2 3 > $OUT/ncu_$1.log 2>&1
       done ;;
  misc) timeout 600 ncu --set full --clock-control none -k regex:'ilsw_policy_act|rb_scatter|rb_gather' -c 14 -f -o $OUT/ncu_misc python tools/ncu_target_misc.py > $OUT/ncu_misc.log 2>&1 ;;
  launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launches.csv python bench.py --steps 2000 --warmup 200 --e2e-steps 20 --no-cpu-baseline --sub none > $OUT/ncu_launches_bench.log 2>&1 ;;
  summarize) python tools/summarize_r2.py $OUT > $OUT/summarize.log 2>&1 ;;
  sanitize) run() { local out=$OUT/sanitize_$1_$2.txt
              timeout 400 compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py $2 3 > $out 2>&1; echo "rc=$?" >> $out
              echo "== $1 $2: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $out | tr '\n' ' ')"; }
            run memcheck sac_ragged; run memcheck sac_hopper_b512_fixed_alpha; run memcheck gail_ragged
            run memcheck td3_ragged; run synccheck sac_hopper_b512_fixed_alpha; run racecheck sac_ragged; run racecheck gail_ragged ;;
esac; done
ls -la $OUT; tail -3 $OUT/replay_bench.txt; grep production $OUT/phase_profile.txt
