#!/bin/bash
# Round-2 GPU call 4: scheduling traces + A/B of occupancy / software-pipelined linearisation builds.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g4_summary.txt
export TFMPC_QUEUE_WTARGET=74 TFMPC_QUEUE_PATIENCE=0
timeout 300 python scripts/queue_trace.py --tag g4_wt74 --streams 8 --rounds 3 > $O/g4_trace1.log 2>&1; tail -n 2 $O/g4_trace1.log
TFMPC_QUEUE_WTARGET=1184 timeout 300 python scripts/queue_trace.py --tag g4_wt1184 --streams 8 --rounds 3 > $O/g4_trace2.log 2>&1; tail -n 2 $O/g4_trace2.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g4_summary.txt; env "$@" > $O/g4_$name.json 2> $O/g4_$name.err; python - "$O/g4_$name.json" >> $O/g4_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    q=d.get("queue_counters") or {}
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f  witer %s rounds %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, q.get("warp_iterations"), q.get("rounds")))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do for v in wps16 wps20 wps12 pl16 pl20 pl12; do
  run ${v}_s8_r$rep TFMPC_B200_LIBDIR=$PWD/ab/$v $B --steps 48 --streams 8
done; done
for v in wps16 pl16 pl12 wps12; do run ${v}_lat TFMPC_QUEUE_WTARGET=1184 TFMPC_B200_LIBDIR=$PWD/ab/$v $B --steps 8 --streams 1; done
cat $O/g4_summary.txt
