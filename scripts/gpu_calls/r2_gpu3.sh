#!/bin/bash
# Round-2 GPU call 3: scheduling sweep (streams x pop-size target x patience) for the queue solver.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g3_summary.txt
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g3_summary.txt; env "$@" > $O/g3_$name.json 2> $O/g3_$name.err; python - "$O/g3_$name.json" >> $O/g3_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    q=d.get("queue_counters") or {}
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f  witer %s rounds %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, q.get("warp_iterations"), q.get("rounds")))
except Exception as e:
    print("FAILED", e)
PY
}
for wt in 37 74 148 296; do for s in 4 8 16; do for pat in 0 4; do
  run wt${wt}_s${s}_p${pat} TFMPC_QUEUE_WTARGET=$wt TFMPC_QUEUE_PATIENCE=$pat $B --steps $((s*6)) --streams $s
done; done; done
run wt1184_s8_p0 TFMPC_QUEUE_PATIENCE=0 $B --steps 48 --streams 8
run wt592_s8_p0 TFMPC_QUEUE_WTARGET=592 TFMPC_QUEUE_PATIENCE=0 $B --steps 48 --streams 8
cat $O/g3_summary.txt
