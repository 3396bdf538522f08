#!/bin/bash
# Round-2 GPU call 23: 10.5 KB per warp (no pad chunk, 32-bit row-offset tables): 18 vs 20 warps, natural order vs rotated chunks.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g23_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q -x > $O/g23_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g23_summary.txt
tail -n 4 $O/g23_pytest_queue.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g23_summary.txt; env "$@" > $O/g23_$name.json 2> $O/g23_$name.err; python - "$O/g23_$name.json" >> $O/g23_summary.txt <<'PY'
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
  run nat18_r$rep TFMPC_QUEUE_WPS=18 $B --steps 48 --streams 8
  run nat20_r$rep TFMPC_X=1 $B --steps 48 --streams 8
  run swz20_r$rep TFMPC_B200_LIBDIR=$PWD/ab/swz $B --steps 48 --streams 8
done
cat $O/g23_summary.txt
