#!/bin/bash
# Round-2 GPU call 11: solo engine -- correctness on the GPU, throughput regression check, lone-batch latency matrix.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g11_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q > $O/g11_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g11_summary.txt
tail -n 12 $O/g11_pytest_queue.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g11_summary.txt; env "$@" > $O/g11_$name.json 2> $O/g11_$name.err; python - "$O/g11_$name.json" >> $O/g11_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    q=d.get("queue_counters") or {}
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f  witer %s rounds %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, q.get("warp_iterations"), q.get("rounds")))
except Exception as e:
    print("FAILED", e)
PY
}
run thr_solo0_wt148 TFMPC_QUEUE_SOLO=0 TFMPC_QUEUE_WTARGET=148 $B --steps 48 --streams 8
run thr_solo4_wt148 TFMPC_QUEUE_SOLO=4 TFMPC_QUEUE_WTARGET=148 $B --steps 48 --streams 8
run thr_solo4_wt296 TFMPC_QUEUE_SOLO=4 TFMPC_QUEUE_WTARGET=296 $B --steps 48 --streams 8
run lat_solo0_wt2664 TFMPC_QUEUE_SOLO=0 TFMPC_QUEUE_WTARGET=2664 $B --steps 8 --streams 1
run lat_solo1_wt2664 TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=2664 $B --steps 8 --streams 1
run lat_solo4_wt2664 TFMPC_QUEUE_SOLO=4 TFMPC_QUEUE_WTARGET=2664 $B --steps 8 --streams 1
run lat_solo8_wt2664 TFMPC_QUEUE_SOLO=8 TFMPC_QUEUE_WTARGET=2664 $B --steps 8 --streams 1
run lat_solo4_wt148 TFMPC_QUEUE_SOLO=4 TFMPC_QUEUE_WTARGET=148 $B --steps 8 --streams 1
run lat_solo4_wt1332 TFMPC_QUEUE_SOLO=4 TFMPC_QUEUE_WTARGET=1332 $B --steps 8 --streams 1
TFMPC_QUEUE_SOLO=4 TFMPC_QUEUE_WTARGET=2664 timeout 300 python scripts/queue_trace.py --tag g11_solo4 --streams 2 --rounds 1 2>&1 | tail -n 2 | tee -a $O/g11_summary.txt
cat $O/g11_summary.txt
