#!/bin/bash
# Round-2 GPU call 26: lane-per-state solver with the nominal read two steps ahead (backward and rollouts): parity tests + bench lines.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g26_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "large or hvac or res or c4 or c5 or mpc or golden" > $O/g26_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g26_summary.txt
tail -n 5 $O/g26_pytest.log
B="timeout 600 python bench.py --no-cpu-baseline --no-clock-sampler --no-extra"
run() { name=$1; shift; echo "== $name" >> $O/g26_summary.txt; env "$@" > $O/g26_$name.json 2> $O/g26_$name.err; python - "$O/g26_$name.json" >> $O/g26_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.2f M/s  ms/step %.3f  seq %.2f ms  e2e %.2f  frac %.3f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"]))
except Exception as e:
    print("FAILED", e)
PY
}
for w in c4 c5s; do
  run ${w}_s1 TFMPC_X=1 $B --workload $w --steps 4 --streams 1
  run ${w}_s4 TFMPC_X=1 $B --workload $w --steps 8 --streams 4
done
run c5 TFMPC_X=1 $B --workload c5 --steps 1 --streams 1
cat $O/g26_summary.txt
