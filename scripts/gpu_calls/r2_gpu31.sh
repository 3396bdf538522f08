#!/bin/bash
# Round-2 GPU call 31: solo engine for lone stragglers while the pipeline drains (throughput mode): A/B at the driver's invocation.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g31_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q > $O/g31_pytest_queue.log 2>&1; echo "pytest_queue rc=$?" | tee -a $O/g31_summary.txt
tail -n 4 $O/g31_pytest_queue.log
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g31_summary.txt; env "$@" > $O/g31_$name.json 2> $O/g31_$name.err; python - "$O/g31_$name.json" >> $O/g31_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  ms/step %.3f  lat %.2f ms  e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  run k20_off_r$rep TFMPC_QUEUE_DRAIN_SOLO=0 $B --steps 20 --warmup 5
  run k20_on_r$rep TFMPC_QUEUE_DRAIN_SOLO=1 $B --steps 20 --warmup 5
  run k64_off_r$rep TFMPC_QUEUE_DRAIN_SOLO=0 $B --steps 64
  run k64_on_r$rep TFMPC_QUEUE_DRAIN_SOLO=1 $B --steps 64
done
cat $O/g31_summary.txt
