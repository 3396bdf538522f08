#!/bin/bash
# Round-2 GPU call 37: latency-mode knobs, second sweep (pop-size target above the number of launched warps).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g37_summary.txt
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong --steps 16 --streams 1"
run() { name=$1; shift; echo "== $name" >> $O/g37_summary.txt; env "$@" > $O/g37_$name.json 2> $O/g37_$name.err; python - "$O/g37_$name.json" >> $O/g37_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("lat %.3f ms" % (d["sequential"]["latency_ms_per_batch"]))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
for wps in 12 13 14; do
  for wt in 2220 2664 3108 3996; do
    run wps${wps}_wt${wt}_r$rep TFMPC_QUEUE_WPS=$wps TFMPC_QUEUE_WTARGET=$wt $B
  done
done
done
paste - - < $O/g37_summary.txt
