#!/bin/bash
# Round-2 GPU call 40: throughput-mode knobs on the final tree (pop-size target, bulk threshold via w_target), K = 64 and K = 20.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g40_summary.txt
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g40_summary.txt; env "$@" > $O/g40_$name.json 2> $O/g40_$name.err; python - "$O/g40_$name.json" >> $O/g40_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  e2e %.1f" % (d["value"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for wt in 74 111 148 185 222 296; do
  run k64_wt$wt TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=$wt $B --steps 64
  run k20_wt$wt TFMPC_QUEUE_MODE=1 TFMPC_QUEUE_WTARGET=$wt $B --steps 20 --warmup 5
done
run k64_auto TFMPC_X=1 $B --steps 64
run k20_auto TFMPC_X=1 $B --steps 20 --warmup 5
paste - - < $O/g40_summary.txt
