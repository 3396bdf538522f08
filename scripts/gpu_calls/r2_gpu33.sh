#!/bin/bash
# Round-2 GPU call 33: ncu evidence refreshed on the final tree: bulk phase and whole solve of k_queue_solve (--set full), DRAM bytes per solve.
mkdir -p gpurun_out
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
TFMPC_QUEUE_MODE=1 timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g33_launches_c3_thr.csv python scripts/profile_solve.py --workload c3 > $O/g33_ncu_c3.log 2>&1
TFMPC_QUEUE_MODE=2 timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g33_launches_c3_lat.csv python scripts/profile_solve.py --workload c3 >> $O/g33_ncu_c3.log 2>&1
for w in c3_thr c3_lat; do echo "== $w"; python scripts/summarize_traffic.py $O/g33_launches_$w.csv; done | tee $O/g33_summary.txt
TFMPC_QUEUE_MODE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g33_queue_bulk python scripts/profile_solve.py --workload c3 --max-iterations 6 > $O/g33_ncu_bulk.log 2>&1
TFMPC_QUEUE_MODE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g33_queue_full python scripts/profile_solve.py --workload c3 > $O/g33_ncu_full.log 2>&1
for n in bulk full; do
  ncu -i $O/g33_queue_$n.ncu-rep --page raw --csv > $O/g33_queue_${n}_raw.csv 2>/dev/null
  ncu -i $O/g33_queue_$n.ncu-rep --page source --csv > $O/g33_queue_${n}_source.csv 2>/dev/null
  python scripts/ncu_summary.py $O/g33_queue_${n}_raw.csv $O/g33_queue_${n}_source.csv > $O/g33_queue_${n}_summary.txt 2>&1
  ncu -i $O/g33_queue_$n.ncu-rep --page details --csv > $O/g33_queue_${n}_details.csv 2>/dev/null
done
head -n 28 $O/g33_queue_bulk_summary.txt
timeout 300 python -m pytest tests/test_gpu_queue.py -q -k auto_mode 2>&1 | tail -n 3 | tee -a $O/g33_summary.txt
