#!/bin/bash
# Round-2 GPU call 18: latency-mode knobs (resident warps per SM, pop-size target, solo threshold) for the lone C3 batch.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g18_summary.txt
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g18_summary.txt; env "$@" > $O/g18_$name.json 2> $O/g18_$name.err; python - "$O/g18_$name.json" >> $O/g18_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
run base TFMPC_X=1 $B --steps 16 --streams 1
run wps16 TFMPC_QUEUE_WPS=16 $B --steps 16 --streams 1
run wps14 TFMPC_QUEUE_WPS=14 $B --steps 16 --streams 1
run wps12 TFMPC_QUEUE_WPS=12 $B --steps 16 --streams 1
run wps16_wt1480 TFMPC_QUEUE_WPS=16 TFMPC_QUEUE_WTARGET=1480 $B --steps 16 --streams 1
run wps14_wt1184 TFMPC_QUEUE_WPS=14 TFMPC_QUEUE_WTARGET=1184 $B --steps 16 --streams 1
run solo2 TFMPC_QUEUE_SOLO=2 $B --steps 16 --streams 1
run ws3200 TFMPC_QUEUE_WSOLO=3200 $B --steps 16 --streams 1
run ws2000 TFMPC_QUEUE_WSOLO=2000 $B --steps 16 --streams 1
run base2 TFMPC_X=1 $B --steps 16 --streams 1
cat $O/g18_summary.txt
