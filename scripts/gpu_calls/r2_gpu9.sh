#!/bin/bash
# Round-2 GPU call 9: half-line staging (11.8 KB smem per warp): correctness + occupancy A/B + bulk ncu.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g9_summary.txt
export TFMPC_QUEUE_WTARGET=148 TFMPC_QUEUE_PATIENCE=0
timeout 400 python -m pytest tests/test_gpu_queue.py -q > $O/g9_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g9_summary.txt
tail -n 3 $O/g9_pytest_queue.log
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g9_summary.txt; env "$@" > $O/g9_$name.json 2> $O/g9_$name.err; python - "$O/g9_$name.json" >> $O/g9_summary.txt <<'PY'
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
  run gs13_r$rep TFMPC_B200_LIBDIR=$PWD/ab/gs13 $B --steps 48 --streams 8
  run hl16_r$rep TFMPC_B200_LIBDIR=$PWD/ab/hl16 $B --steps 48 --streams 8
  run hl18_r$rep TFMPC_QUEUE_WPS=18 TFMPC_B200_LIBDIR=$PWD/ab/hl18 $B --steps 48 --streams 8
  run hl20_w18_r$rep TFMPC_QUEUE_WPS=18 TFMPC_B200_LIBDIR=$PWD/ab/hl20 $B --steps 48 --streams 8
  run hl18_w14_r$rep TFMPC_QUEUE_WPS=14 TFMPC_B200_LIBDIR=$PWD/ab/hl18 $B --steps 48 --streams 8
done
run hl18_s16 TFMPC_QUEUE_WPS=18 TFMPC_B200_LIBDIR=$PWD/ab/hl18 $B --steps 96 --streams 16
run hl18_lat1184 TFMPC_QUEUE_WPS=18 TFMPC_QUEUE_WTARGET=1184 TFMPC_B200_LIBDIR=$PWD/ab/hl18 $B --steps 8 --streams 1
TFMPC_QUEUE_WPS=18 TFMPC_B200_LIBDIR=$PWD/ab/hl18 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g9_bulk_hl18 python scripts/profile_solve.py --workload c3 --max-iterations 6 > $O/g9_ncu_hl18.log 2>&1
cat $O/g9_summary.txt
