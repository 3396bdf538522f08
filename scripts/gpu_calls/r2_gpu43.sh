#!/bin/bash
# Round-2 GPU call 43: throughput-mode solves that find no other solve in its bulk phase spread out like a lone batch: A/B.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g43_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q > $O/g43_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g43_summary.txt
tail -n 3 $O/g43_pytest.log
B="timeout 240 python bench.py --gpus 1 --no-cpu-baseline --no-clock-sampler --no-extra --no-strong"
run() { name=$1; shift; echo "== $name" >> $O/g43_summary.txt; env "$@" > $O/g43_$name.json 2> $O/g43_$name.err; python - "$O/g43_$name.json" >> $O/g43_summary.txt <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f M/s  e2e %.1f  lat %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["sequential"]["latency_ms_per_batch"]))
except Exception as e:
    print("FAILED", e)
PY
}
for rep in 1 2; do
  run k20_off_r$rep TFMPC_QUEUE_DRAIN_SOLO=0 $B --steps 20 --warmup 5
  run k20_on_r$rep TFMPC_X=1 $B --steps 20 --warmup 5
  run k64_off_r$rep TFMPC_QUEUE_DRAIN_SOLO=0 $B --steps 64
  run k64_on_r$rep TFMPC_X=1 $B --steps 64
done
paste - - < $O/g43_summary.txt
