#!/bin/bash
# Round-2 GPU call 25: lane-per-state solver with the nominal actions and k staged in shared memory: parity tests, then A/B
# against global reads (TFMPC_KW_NO_STAGING=1) on C4, C5 (single solve) and C5 (MPC loop).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g25_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "large or hvac or res or c4 or c5 or mpc or golden" > $O/g25_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g25_summary.txt
tail -n 5 $O/g25_pytest.log
B="timeout 600 python bench.py --no-cpu-baseline --no-clock-sampler --no-extra"
run() { name=$1; shift; echo "== $name" >> $O/g25_summary.txt; env "$@" > $O/g25_$name.json 2> $O/g25_$name.err; python - "$O/g25_$name.json" >> $O/g25_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.2f M/s  ms/step %.3f  seq %.2f ms  e2e %.2f  frac %.3f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"]))
except Exception as e:
    print("FAILED", e)
PY
}
for w in c4 c5s; do
  run ${w}_global TFMPC_KW_NO_STAGING=1 $B --workload $w --steps 4 --streams 1
  run ${w}_staged TFMPC_X=1 $B --workload $w --steps 4 --streams 1
  run ${w}_staged_s4 TFMPC_X=1 $B --workload $w --steps 8 --streams 4
done
run c5_global TFMPC_KW_NO_STAGING=1 $B --workload c5 --steps 1 --streams 1
run c5_staged TFMPC_X=1 $B --workload c5 --steps 1 --streams 1
cat $O/g25_summary.txt
