#!/bin/bash
# Round-2 GPU call 12: ncu source-level capture of the solo engine alone (148 lone problems, one warp each).
O=gpurun_out
TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=1048576 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_queue_solve -s 1 -c 1 -f -o $O/g12_solo python scripts/profile_solve.py --workload c3 --batch 148 > $O/g12_ncu.log 2>&1
tail -n 3 $O/g12_ncu.log
ncu -i $O/g12_solo.ncu-rep --page raw --csv > $O/g12_raw.csv 2>/dev/null
ncu -i $O/g12_solo.ncu-rep --page source --csv > $O/g12_source.csv 2>/dev/null
ls -la $O/g12_*
