#!/bin/bash
# Round-2 GPU call 46: experiment -- the last N batches of a pipelined run in latency mode (nothing behind them wants the SMs).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g46_summary.txt
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g46_summary.txt; env "$@" > $O/g46_$name.json 2> $O/g46_$name.err; python - "$O/g46_$name.json" >> $O/g46_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  for n in 0 1 2 4 8; do
    run k20_tail${n}_r$rep TFMPC_BENCH_TAIL_LATENCY=$n $B --steps 20 --warmup 5
  done
done
run k64_tail4 TFMPC_BENCH_TAIL_LATENCY=4 $B --steps 64
run k64_tail0 TFMPC_BENCH_TAIL_LATENCY=0 $B --steps 64
paste - - < $O/g46_summary.txt
