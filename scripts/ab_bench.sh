#!/bin/bash
# A/B harness: run bench.py against several builds of the library (ab/<name>/libtfmpc_b200.so), interleaved, REPS times.
# usage: scripts/ab_bench.sh "A B C" [reps] [steps]
VARIANTS=${1:-"A B"}; REPS=${2:-3}; STEPS=${3:-16}
mkdir -p gpurun_out
for r in $(seq 1 $REPS); do
  for v in $VARIANTS; do
    TFMPC_B200_LIBDIR=$PWD/ab/$v python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v rep$r pipelined %.1fM/s %.3f ms/step | sequential %.3f ms | e2e %.1fM/s' % (d['value']/1e6, d['ms_per_step'], d['sequential']['latency_ms_per_batch'], d['e2e']['value']/1e6))"
  done
done | tee gpurun_out/ab_bench.txt
