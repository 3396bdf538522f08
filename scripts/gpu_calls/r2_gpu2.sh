#!/bin/bash
# Round-2 GPU call 2: queue solver with the semaphore acquire; schedule sweeps; launch list; one full capture.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g2_summary.txt
timeout 400 python -m pytest tests/test_gpu_queue.py -x -q -s > $O/g2_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g2_summary.txt
tail -n 3 $O/g2_pytest_queue.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g2_summary.txt; env "$@" > $O/g2_$name.json 2> $O/g2_$name.err; python - "$O/g2_$name.json" >> $O/g2_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  seq %.1f M/s lat %.2f ms  e2e %.1f  launches %s  status %s  q %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["value"]/1e6, d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d.get("gpu_launches"), d["config"].get("status_histogram"), d.get("queue_counters")))
except Exception as e:
    print("FAILED", e)
PY
}
run queue_s1   $B --steps 8 --streams 1
run queue_s2   $B --steps 16 --streams 2
run queue_s4   $B --steps 32 --streams 4
run queue_s8   $B --steps 32 --streams 8
run queue_newton_s4 TFMPC_QP=newton $B --steps 16 --streams 4
for wt in 148 296 592 2368; do run queue_wt${wt}_s4 TFMPC_QUEUE_WTARGET=$wt $B --steps 32 --streams 4; run queue_wt${wt}_s1 TFMPC_QUEUE_WTARGET=$wt $B --steps 8 --streams 1; done
run queue_wps12_s4 TFMPC_QUEUE_WPS=12 $B --steps 32 --streams 4
run queue_wps8_s4 TFMPC_QUEUE_WPS=8 $B --steps 32 --streams 4
run queue_wps12_s1 TFMPC_QUEUE_WPS=12 $B --steps 8 --streams 1
run queue_pat0_s4 TFMPC_QUEUE_PATIENCE=0 $B --steps 32 --streams 4
run queue_pat16_s4 TFMPC_QUEUE_PATIENCE=16 $B --steps 32 --streams 4
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g2_launches_queue.csv python scripts/profile_solve.py --workload c3 > $O/g2_ncu1.log 2>&1
tail -n 2 $O/g2_ncu1.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g2_full_queue python scripts/profile_solve.py --workload c3 > $O/g2_ncu2.log 2>&1
tail -n 2 $O/g2_ncu2.log
cat $O/g2_summary.txt
