#!/bin/bash
# Round-2 GPU call 34: explicit fma chains in both Q-assembly shapes (schedule-independent fp32 results?), sticky solo that yields
# when tickets are queued: queue tests, throughput mode with the solo engine on against off, latency.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g34_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q > $O/g34_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g34_summary.txt
tail -n 12 $O/g34_pytest_queue.log
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g34_summary.txt; env "$@" > $O/g34_$name.json 2> $O/g34_$name.err; python - "$O/g34_$name.json" >> $O/g34_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  run k64_default_r$rep TFMPC_X=1 $B --steps 64
  run k64_solo1_r$rep TFMPC_QUEUE_SOLO=1 $B --steps 64
  run k20_default_r$rep TFMPC_X=1 $B --steps 20 --warmup 5
  run k20_solo1_r$rep TFMPC_QUEUE_SOLO=1 $B --steps 20 --warmup 5
done
cat $O/g34_summary.txt
