#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
export TFMPC_QUEUE_WTARGET=74 TFMPC_QUEUE_PATIENCE=0
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g5_full_bulk python scripts/profile_solve.py --workload c3 --max-iterations 6 > $O/g5_ncu.log 2>&1
tail -n 3 $O/g5_ncu.log
