#!/bin/bash
# Round-2 GPU call 36: latency-mode knobs on the final tree (resident warps per SM x pop-size target), lone C3 batch.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g36_summary.txt
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong --steps 16 --streams 1"
run() { name=$1; shift; echo "== $name" >> $O/g36_summary.txt; env "$@" > $O/g36_$name.json 2> $O/g36_$name.err; python - "$O/g36_$name.json" >> $O/g36_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("lat %.3f ms" % (d["sequential"]["latency_ms_per_batch"]))
except Exception as e:
    print("FAILED", e)
PY
}
for wps in 12 13 14 15; do
  for wt in 1332 1776 2220; do
    run wps${wps}_wt${wt} TFMPC_QUEUE_WPS=$wps TFMPC_QUEUE_WTARGET=$wt $B
  done
done
run base TFMPC_X=1 $B
run base2 TFMPC_X=1 $B
paste - - < $O/g36_summary.txt
