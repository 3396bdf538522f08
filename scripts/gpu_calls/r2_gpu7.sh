#!/bin/bash
# Round-2 GPU call 7: gains staged through shared memory (cp.async half-line ring): correctness + A/B + bulk ncu.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g7_summary.txt
export TFMPC_QUEUE_WTARGET=74 TFMPC_QUEUE_PATIENCE=0
timeout 400 python -m pytest tests/test_gpu_queue.py -x -q > $O/g7_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g7_summary.txt
tail -n 3 $O/g7_pytest_queue.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g7_summary.txt; env "$@" > $O/g7_$name.json 2> $O/g7_$name.err; python - "$O/g7_$name.json" >> $O/g7_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    q=d.get("queue_counters") or {}
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f  witer %s rounds %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, q.get("warp_iterations"), q.get("rounds")))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  run base16_r$rep TFMPC_B200_LIBDIR=$PWD/ab/wps16 $B --steps 48 --streams 8
  for w in 12 10 8; do run gs12_w${w}_r$rep TFMPC_QUEUE_WPS=$w TFMPC_B200_LIBDIR=$PWD/ab/gs12 $B --steps 48 --streams 8; done
  run gs13_r$rep TFMPC_B200_LIBDIR=$PWD/ab/gs13 $B --steps 48 --streams 8
done
run gs12_lat1184 TFMPC_QUEUE_WTARGET=1184 TFMPC_B200_LIBDIR=$PWD/ab/gs12 $B --steps 8 --streams 1
run gs12_lat2368 TFMPC_QUEUE_WTARGET=2368 TFMPC_B200_LIBDIR=$PWD/ab/gs12 $B --steps 8 --streams 1
TFMPC_B200_LIBDIR=$PWD/ab/gs12 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g7_bulk_gs12 python scripts/profile_solve.py --workload c3 --max-iterations 6 > $O/g7_ncu_gs12.log 2>&1
cat $O/g7_summary.txt
