#!/bin/bash
# Round-2 GPU call 14: auto scheduling mode (latency when alone, throughput when pipelined) + lane-parallel solo value update.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g14_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q > $O/g14_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g14_summary.txt
tail -n 8 $O/g14_pytest_queue.log
python scripts/solo_micro.py 2>&1 | tail -n 2 | tee -a $O/g14_summary.txt
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g14_summary.txt; env "$@" > $O/g14_$name.json 2> $O/g14_$name.err; python - "$O/g14_$name.json" >> $O/g14_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    q=d.get("queue_counters") or {}
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f  witer %s rounds %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, q.get("warp_iterations"), q.get("rounds")))
except Exception as e:
    print("FAILED", e)
PY
}
run auto_s8 TFMPC_X=1 $B --steps 48 --streams 8
run auto_s1 TFMPC_X=1 $B --steps 8 --streams 1
run auto_s2 TFMPC_X=1 $B --steps 16 --streams 2
run thr_s8 TFMPC_QUEUE_MODE=1 $B --steps 48 --streams 8
run lat_wt1480 TFMPC_QUEUE_MODE=2 TFMPC_QUEUE_WTARGET=1480 $B --steps 8 --streams 1
run lat_wt2072 TFMPC_QUEUE_MODE=2 TFMPC_QUEUE_WTARGET=2072 $B --steps 8 --streams 1
cat $O/g14_summary.txt
