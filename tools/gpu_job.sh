#!/bin/bash
# development job for one gpurun call: tile check, GPU tests, phase profiles.  Usage: tools/gpu_job.sh <tag> [steps...]
TAG=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    tc5) timeout 200 ./tools/tc5_test > gpurun_out/${TAG}_tc5.txt 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_tc5.txt ;;
    tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_tests.log ;;
    tests_all) timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_tests.log ;;
    pp_td3) timeout 300 python tools/phase_profile.py td3_humanoid > gpurun_out/${TAG}_pp_td3.txt 2>&1 ;;
    pp_her) timeout 300 python tools/phase_profile.py her_td3_pick > gpurun_out/${TAG}_pp_her.txt 2>&1 ;;
    pp_sac) timeout 300 python tools/phase_profile.py sac_hopper > gpurun_out/${TAG}_pp_sac.txt 2>&1 ;;
    pp_gail) timeout 300 python tools/phase_profile.py gail_walker > gpurun_out/${TAG}_pp_gail.txt 2>&1 ;;
    pp_ant) timeout 300 python tools/phase_profile.py sac_ant > gpurun_out/${TAG}_pp_ant.txt 2>&1 ;;
    bench) timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ;;
    benchlong) timeout 600 python bench.py > gpurun_out/${TAG}_benchlong.json 2> gpurun_out/${TAG}_benchlong.err ;;
    pp_sac_stamps) timeout 300 python tools/phase_profile.py sac_hopper gail_walker > gpurun_out/${TAG}_pp_sac_stamps.txt 2>&1 ;;
    ncu_td3) timeout 600 ncu --set full --clock-control none --import-source on -k regex:ilsw_engine_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_td3 python tools/ncu_target.py td3_humanoid 4 3 > gpurun_out/${TAG}_ncu_td3.log 2>&1 ;;
    ncu_sac) timeout 600 ncu --set full --clock-control none --import-source on -k regex:ilsw_engine_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_sac python tools/ncu_target.py sac_hopper 10 3 > gpurun_out/${TAG}_ncu_sac.log 2>&1 ;;
    sanitize) timeout 3000 tools/sanitize.sh ${TAG} > gpurun_out/${TAG}_sanitize_summary.txt 2>&1 ;;
    ab_sac) timeout 600 tools/ab.sh "sac_hopper --no-tile-stamps" variants/r1.so ilswiss_b200/csrc/libilswiss_b200.so variants/r1.so ilswiss_b200/csrc/libilswiss_b200.so > gpurun_out/${TAG}_ab_sac.txt 2>&1 ;;
    smi) nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/${TAG}_smi.txt 2>&1 ;;
    *) echo "unknown step $step" ;;
  esac
done
for f in gpurun_out/${TAG}_*; do echo "== $f"; tail -c 1500 $f; echo; done
