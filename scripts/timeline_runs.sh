#!/bin/bash
# run bench.py N times and keep every JSON line (to catch the slow pipelined mode): scripts/timeline_runs.sh N steps [args]
N=$1; STEPS=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/timeline_runs.jsonl
for r in $(seq 1 $N); do
  python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline "$@" 2>/dev/null >> gpurun_out/timeline_runs.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/timeline_runs.jsonl"):
    d = json.loads(l)
    print("%.1fM/s %.3f ms/step" % (d["value"]/1e6, d["ms_per_step"]))
PY
