#!/bin/bash
# Round-end measurement on ONE B200 box (run through gpurun): bench lines for every workload, the reference arm,
# the ncu launch list + one full capture of the engine kernel, and the per-phase profile.  Outputs: gpurun_out/final/
set -u
OUT=gpurun_out/final; mkdir -p $OUT
python bench.py > $OUT/bench_sac_hopper.json 2> $OUT/bench_sac_hopper.err
for w in gail_walker td3_humanoid sac_ant her_td3_pick; do
  python bench.py --workload $w --steps 10000 --warmup 1000 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
python bench.py --impl reference --steps 1500 --warmup 20 > $OUT/bench_reference_sac_hopper.json 2> $OUT/bench_reference.err
python tools/phase_profile.py sac_hopper gail_walker td3_humanoid > $OUT/phase_profile.txt 2>&1
if [ "${SKIP_NCU:-0}" = "1" ]; then ls $OUT; for f in $OUT/bench_*.json; do python -c "
import json, sys
d = json.load(open('$f')); print('$f'.split('/')[-1], d.get('value'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))"; done; exit 0; fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launches.csv \
    python bench.py --steps 3000 --warmup 1000 --e2e-steps 20 --no-cpu-baseline --precision 3 > $OUT/ncu_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ilsw_engine_kernel -s 2 -c 1 -f -o $OUT/engine_full \
    python tools/ncu_target.py sac_hopper 50 3 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/engine_full.ncu-rep --page raw --csv > $OUT/engine_full_raw.csv 2>/dev/null
ls -la $OUT
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
