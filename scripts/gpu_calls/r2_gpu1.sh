#!/bin/bash
# Round-2 GPU call 1: correctness of the queue solver on hardware, first A/B of schedules, first launch list.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/g1_smi.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_queue.py -x -q -s > $O/g1_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g1_summary.txt
tail -n 5 $O/g1_pytest_queue.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q > $O/g1_pytest_parity.log 2>&1; echo "pytest_parity rc=$?" | tee -a $O/g1_summary.txt
tail -n 8 $O/g1_pytest_parity.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g1_summary.txt; env "$@" > $O/g1_$name.json 2> $O/g1_$name.err; python - "$O/g1_$name.json" >> $O/g1_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  seq %.1f M/s lat %.2f ms  e2e %.1f  launches %s  status %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["value"]/1e6, d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d.get("gpu_launches"), d["config"].get("status_histogram")))
except Exception as e:
    print("FAILED", e)
PY
}
run ticks_s8   TFMPC_SOLVER=ticks $B --steps 32
run queue_s8   $B --steps 32
run queue_s1   $B --steps 8 --streams 1
run queue_s2   $B --steps 16 --streams 2
run queue_s3   $B --steps 24 --streams 3
run queue_s4   $B --steps 32 --streams 4
run queue_newton_s4 TFMPC_QP=newton $B --steps 16 --streams 4
for wt in 296 592 2368; do run queue_wt${wt}_s4 TFMPC_QUEUE_WTARGET=$wt $B --steps 32 --streams 4; run queue_wt${wt}_s1 TFMPC_QUEUE_WTARGET=$wt $B --steps 8 --streams 1; done
run queue_wps12_s4 TFMPC_QUEUE_WPS=12 $B --steps 32 --streams 4
run queue_wps8_s4 TFMPC_QUEUE_WPS=8 $B --steps 32 --streams 4
run queue_pat0_s4 TFMPC_QUEUE_PATIENCE=0 $B --steps 32 --streams 4
run queue_pat16_s4 TFMPC_QUEUE_PATIENCE=16 $B --steps 32 --streams 4
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/g1_launches_queue.csv python scripts/profile_solve.py --workload c3 > $O/g1_ncu1.log 2>&1
tail -n 3 $O/g1_ncu1.log
cat $O/g1_summary.txt
