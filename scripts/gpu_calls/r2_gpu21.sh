#!/bin/bash
# Round-2 GPU call 21: A/B linearisation one step ahead in the lane-per-problem backward (interleaved runs, same box).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g21_summary.txt
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g21_summary.txt; env "$@" > $O/g21_$name.json 2> $O/g21_$name.err; python - "$O/g21_$name.json" >> $O/g21_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2 3; do
  run nolin_r$rep TFMPC_B200_LIBDIR=$PWD/ab/nolinahead $B --steps 48 --streams 8
  run lin_r$rep TFMPC_B200_LIBDIR=$PWD/ab/linahead $B --steps 48 --streams 8
done
cat $O/g21_summary.txt
