#!/bin/bash
# Round-2 GPU call 35: whole GPU suite after the explicit-fma Q assembly (schedule-independent fp32 results) + latency / driver lines.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g35_summary.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > $O/g35_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g35_summary.txt
tail -n 8 $O/g35_pytest.log | tee -a $O/g35_summary.txt
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra"
run() { name=$1; shift; echo "== $name" >> $O/g35_summary.txt; env "$@" > $O/g35_$name.json 2> $O/g35_$name.err; python - "$O/g35_$name.json" >> $O/g35_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f strong %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["extra"]["strong"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
run k64_r1 TFMPC_X=1 $B --steps 64
run k20_r1 TFMPC_X=1 $B --steps 20 --warmup 5
run k64_r2 TFMPC_X=1 $B --steps 64
cat $O/g35_summary.txt | tail -8
