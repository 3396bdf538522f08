#!/bin/bash
# Round-2 GPU call 16: the reworked bench.py (default line), DRAM traffic of one solve per workload (ncu), ncu --set full of the
# queue kernel over a whole solve, re-run of the tests fixed since call 15.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "retry" > $O/g16_pytest_retry.log 2>&1; echo "pytest retry rc=$?" | tee $O/g16_summary.txt
tail -n 4 $O/g16_pytest_retry.log
( time timeout 900 python bench.py ) > $O/g16_bench_default.json 2> $O/g16_bench_default.err; echo "bench rc=$?" | tee -a $O/g16_summary.txt
tail -n 5 $O/g16_bench_default.err
head -c 1500 $O/g16_bench_default.json; echo
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
TFMPC_QUEUE_MODE=1 timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g16_launches_c3_thr.csv python scripts/profile_solve.py --workload c3 > $O/g16_ncu_c3.log 2>&1
TFMPC_QUEUE_MODE=2 timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g16_launches_c3_lat.csv python scripts/profile_solve.py --workload c3 >> $O/g16_ncu_c3.log 2>&1
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g16_launches_c4.csv python scripts/profile_solve.py --workload c4 > $O/g16_ncu_c4.log 2>&1
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g16_launches_c5s.csv python scripts/profile_solve.py --workload c5s > $O/g16_ncu_c5s.log 2>&1
for w in c3_thr c3_lat c4 c5s; do echo "== $w"; python scripts/summarize_traffic.py $O/g16_launches_$w.csv; done | tee -a $O/g16_summary.txt
TFMPC_QUEUE_MODE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g16_queue_full python scripts/profile_solve.py --workload c3 > $O/g16_ncu_full.log 2>&1
ncu -i $O/g16_queue_full.ncu-rep --page raw --csv > $O/g16_queue_full_raw.csv 2>/dev/null
ncu -i $O/g16_queue_full.ncu-rep --page source --csv > $O/g16_queue_full_source.csv 2>/dev/null
python scripts/ncu_summary.py $O/g16_queue_full_raw.csv $O/g16_queue_full_source.csv > $O/g16_queue_full_summary.txt 2>&1
head -n 30 $O/g16_queue_full_summary.txt
