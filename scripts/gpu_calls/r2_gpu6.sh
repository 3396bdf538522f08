#!/bin/bash
# Round-2 GPU call 6: A/B gain prefetch distance / L2 hints / warps per SM; bulk-phase ncu of the hinted build.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g6_summary.txt
export TFMPC_QUEUE_WTARGET=74 TFMPC_QUEUE_PATIENCE=0
B="timeout 240 python bench.py --no-cpu-baseline --no-clock-sampler"
run() { name=$1; shift; echo "== $name" >> $O/g6_summary.txt; env "$@" > $O/g6_$name.json 2> $O/g6_$name.err; python - "$O/g6_$name.json" >> $O/g6_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    q=d.get("queue_counters") or {}
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f  witer %s rounds %s" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, q.get("warp_iterations"), q.get("rounds")))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do for v in wps16 pf3 pf3h; do for w in 16 12 8; do
  run ${v}_w${w}_r$rep TFMPC_QUEUE_WPS=$w TFMPC_B200_LIBDIR=$PWD/ab/$v $B --steps 48 --streams 8
done; done; done
for v in pf3 pf3h; do
TFMPC_B200_LIBDIR=$PWD/ab/$v timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_queue_solve -c 1 -f -o $O/g6_bulk_$v python scripts/profile_solve.py --workload c3 --max-iterations 6 > $O/g6_ncu_$v.log 2>&1
done
cat $O/g6_summary.txt
