#!/bin/bash
# interleaved bench.py runs under (env, args) variants "ENV=..|--args": scripts/ab_matrix.sh reps steps "X=0|--flag" ...
REPS=$1; STEPS=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/ab_matrix.jsonl
for r in $(seq 1 $REPS); do
  for v in "$@"; do
    e=${v%%|*}; a=${v#*|}
    env $e python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline $a 2>/dev/null | python -c "
import json,sys
l=sys.stdin.read(); d=json.loads(l)
open('gpurun_out/ab_matrix.jsonl','a').write(json.dumps({'variant':'$v','rep':$r,'line':d})+chr(10))
tl=d['pipeline_timeline_ms']['steps']; extra=''
if tl:
    first_end=min(b for _,a,b in tl); extra=' | solves started before first finished: %d, first solve %.1f ms' % (sum(1 for _,a,b in tl if a<first_end), tl[0][2]-tl[0][1])
print('[$v] rep$r pipelined %.1fM/s %.3f ms/step | seq %.3f ms | e2e %.1fM/s%s' % (d['value']/1e6, d['ms_per_step'], d['sequential']['latency_ms_per_batch'], d['e2e']['value']/1e6, extra))"
  done
done | tee gpurun_out/ab_matrix.txt
