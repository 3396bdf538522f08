#!/bin/bash
# Round-2 GPU call 29: streams sweep of the default bench (pipelined depth), then the whole GPU suite on the final tree.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g29_summary.txt
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g29_summary.txt; env "$@" > $O/g29_$name.json 2> $O/g29_$name.err; python - "$O/g29_$name.json" >> $O/g29_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  run s8_r$rep TFMPC_X=1 $B --steps 64 --streams 8
  run s12_r$rep TFMPC_X=1 $B --steps 96 --streams 12
  run s16_r$rep TFMPC_X=1 $B --steps 128 --streams 16
done
run s16_64 TFMPC_X=1 $B --steps 64 --streams 16
run s12_64 TFMPC_X=1 $B --steps 64 --streams 12
cat $O/g29_summary.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > $O/g29_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g29_summary.txt
tail -n 8 $O/g29_pytest.log
