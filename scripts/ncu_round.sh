#!/bin/bash
# The ncu passes whose summaries are committed under profiles/ (run on the GPU box, one GPU):
#   1. launch list of ONE C3 solve: duration + DRAM bytes per launch, default cache control (cold L2 per launch)
#   2. the same with --cache-control none (L2 state as in a real run: what one launch leaves, the next finds)
#   3. --set full of a heavy k_tick_backward / k_tick_linesearch launch (tick 2) and of a tail backward launch
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_cold.csv python scripts/profile_solve.py --workload c3 > gpurun_out/ncu1.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/launches_warm.csv python scripts/profile_solve.py --workload c3 > gpurun_out/ncu2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_tick_backward -s 2 -c 1 -f -o gpurun_out/full_bwd_heavy python scripts/profile_solve.py --workload c3 > gpurun_out/ncu3.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_tick_linesearch -s 2 -c 1 -f -o gpurun_out/full_ls_heavy python scripts/profile_solve.py --workload c3 > gpurun_out/ncu4.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_tick_backward -s 60 -c 1 -f -o gpurun_out/full_bwd_tail python scripts/profile_solve.py --workload c3 > gpurun_out/ncu5.log 2>&1
tail -n 2 gpurun_out/ncu1.log; tail -n 2 gpurun_out/ncu3.log
