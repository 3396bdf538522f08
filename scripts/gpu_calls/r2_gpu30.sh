#!/bin/bash
# Round-2 GPU call 30: the driver's invocation (--steps 20 --warmup 5): pipeline depth for a short timed region.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g30_summary.txt
B="timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g30_summary.txt; env "$@" > $O/g30_$name.json 2> $O/g30_$name.err; python - "$O/g30_$name.json" >> $O/g30_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  for s in 4 5 6 8 10; do
    run s${s}_r$rep TFMPC_X=1 $B --streams $s
  done
done
cat $O/g30_summary.txt
