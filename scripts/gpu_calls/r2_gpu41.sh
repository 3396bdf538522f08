#!/bin/bash
# Round-2 GPU call 41: throughput mode, smaller pop-size targets (more problems per warp while a batch drains), K = 64 and K = 20.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g41_summary.txt
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g41_summary.txt; env "$@" > $O/g41_$name.json 2> $O/g41_$name.err; python - "$O/g41_$name.json" >> $O/g41_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  e2e %.1f" % (d["value"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for wt in 9 18 37 55 74; do
  run k64_wt$wt TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=$wt $B --steps 64
  run k20_wt$wt TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=$wt $B --steps 20 --warmup 5
done
run k64_wt37_nodrain TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=37 TFMPC_QUEUE_DRAIN_SOLO=0 $B --steps 64
run k20_wt37_nodrain TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=37 TFMPC_QUEUE_DRAIN_SOLO=0 $B --steps 20 --warmup 5
run k64_wt37_s12 TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=37 $B --steps 64 --streams 12
run k20_wt37_s12 TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=37 $B --steps 20 --warmup 5 --streams 12
paste - - < $O/g41_summary.txt
