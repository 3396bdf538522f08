#!/bin/bash
# Round-2 GPU call 22: 10 KB of shared memory per warp (swizzled staging rows, row keys by shuffle, solo search in passes):
# correctness, then 18 vs 20 resident warps per SM against the previous build (interleaved runs, same box).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g22_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q > $O/g22_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g22_summary.txt
tail -n 6 $O/g22_pytest_queue.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g22_summary.txt; env "$@" > $O/g22_$name.json 2> $O/g22_$name.err; python - "$O/g22_$name.json" >> $O/g22_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  run old_r$rep TFMPC_B200_LIBDIR=$PWD/ab/old $B --steps 48 --streams 8
  run new18_r$rep TFMPC_QUEUE_WPS=18 $B --steps 48 --streams 8
  run new20_r$rep TFMPC_X=1 $B --steps 48 --streams 8
done
run new20_s12 TFMPC_X=1 $B --steps 72 --streams 12
run new20_s16 TFMPC_X=1 $B --steps 96 --streams 16
cat $O/g22_summary.txt
