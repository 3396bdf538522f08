#!/bin/bash
# Round-2 GPU call 44: ncu --set full of the lane-per-state solver (kw_solve) on C4 and C5 (single solve): what bounds it.
mkdir -p gpurun_out
O=gpurun_out
for w in c4 c5s; do
  timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:kw_solve -c 1 -f -o $O/g44_kw_$w python scripts/profile_solve.py --workload $w --batch 8192 > $O/g44_ncu_$w.log 2>&1
  ncu -i $O/g44_kw_$w.ncu-rep --page raw --csv > $O/g44_kw_${w}_raw.csv 2>/dev/null
  ncu -i $O/g44_kw_$w.ncu-rep --page source --csv > $O/g44_kw_${w}_source.csv 2>/dev/null
  python scripts/ncu_summary.py $O/g44_kw_${w}_raw.csv $O/g44_kw_${w}_source.csv > $O/g44_kw_${w}_summary.txt 2>&1
  echo "== $w"; head -n 30 $O/g44_kw_${w}_summary.txt
done
