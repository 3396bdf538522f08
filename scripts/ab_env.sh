#!/bin/bash
# compare bench.py under different environment settings, interleaved: scripts/ab_env.sh reps steps "VAR=1" "VAR=2" ...
REPS=$1; STEPS=$2; shift 2
mkdir -p gpurun_out
for r in $(seq 1 $REPS); do
  for v in "$@"; do
    env $v python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v rep$r pipelined %.1fM/s %.3f ms/step | sequential %.3f ms | e2e %.1fM/s' % (d['value']/1e6, d['ms_per_step'], d['sequential']['latency_ms_per_batch'], d['e2e']['value']/1e6))"
  done
done | tee gpurun_out/ab_env.txt
