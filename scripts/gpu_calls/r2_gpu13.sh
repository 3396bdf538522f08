#!/bin/bash
# Round-2 GPU call 13: lone-batch latency against (pop-size target, solo threshold, solo_max); throughput check with solo on.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g13_summary.txt
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g13_summary.txt; env "$@" > $O/g13_$name.json 2> $O/g13_$name.err; python - "$O/g13_$name.json" >> $O/g13_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    q=d.get("queue_counters") or {}
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f  witer %s rounds %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, q.get("warp_iterations"), q.get("rounds")))
except Exception as e:
    print("FAILED", e)
PY
}
run thr_solo1_wt148 TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=148 $B --steps 48 --streams 8
run thr_solo1_wt148_ws592 TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=148 TFMPC_QUEUE_WSOLO=592 $B --steps 48 --streams 8
run thr_solo0_wt148 TFMPC_QUEUE_SOLO=0 TFMPC_QUEUE_WTARGET=148 $B --steps 48 --streams 8
for wt in 592 888 1184 1776 2664; do
  run lat_solo1_wt${wt}_ws2664 TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=$wt TFMPC_QUEUE_WSOLO=2664 $B --steps 6 --streams 1
done
run lat_solo1_wt1184_ws1776 TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=1184 TFMPC_QUEUE_WSOLO=1776 $B --steps 6 --streams 1
run lat_solo1_wt1184_ws1184 TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=1184 TFMPC_QUEUE_WSOLO=1184 $B --steps 6 --streams 1
run lat_solo2_wt1184_ws2664 TFMPC_QUEUE_SOLO=2 TFMPC_QUEUE_WTARGET=1184 TFMPC_QUEUE_WSOLO=2664 $B --steps 6 --streams 1
run lat_solo1_wt148_ws2664 TFMPC_QUEUE_SOLO=1 TFMPC_QUEUE_WTARGET=148 TFMPC_QUEUE_WSOLO=2664 $B --steps 6 --streams 1
cat $O/g13_summary.txt
