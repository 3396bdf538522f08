#!/bin/bash
# Round-2 GPU call 48: repeatability of the driver invocation with and without the last-batch rule (full line, no extras).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g48_summary.txt
for rep in 1 2 3; do
  for t in 1 0; do
    TFMPC_BENCH_TAIL_LATENCY=$t timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra > $O/g48_t${t}_r$rep.json 2> $O/g48_t${t}_r$rep.err
    python - "$O/g48_t${t}_r$rep.json" "t$t r$rep" >> $O/g48_summary.txt <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "value %.1f M/s ms/step %.3f e2e %.1f lat %.2f" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["sequential"]["latency_ms_per_batch"]))
PY
  done
done
cat $O/g48_summary.txt
