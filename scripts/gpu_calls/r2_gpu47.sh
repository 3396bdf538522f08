#!/bin/bash
# Round-2 GPU call 47: bench lines after the last-batch rule (driver invocation, default), 1 GPU.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/g47_bench_driver.json 2> $O/g47_bench_driver.err; echo "bench driver rc=$?" | tee $O/g47_summary.txt
( time timeout 900 python bench.py ) > $O/g47_bench_default.json 2> $O/g47_bench_default.err; echo "bench default rc=$?" | tee -a $O/g47_summary.txt
python - <<'PY' | tee -a gpurun_out/g47_summary.txt
import json
for f in ("driver","default"):
    d=json.loads(open(f"gpurun_out/g47_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.1f M/s ms/step %.3f lat %.2f ms e2e %.1f frac %.3f launches %d parity same %.4f strong %.1f" % (
        d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["gpu_launches"], d["parity"]["same_iterations"], d["extra"]["strong"]["value"]/1e6))
    for k,v in d["extra"]["workloads"].items():
        print("   ", k, "value %.4g %s" % (v.get("value",0), v.get("unit")), "frac", round((v.get("roofline") or {}).get("frac") or 0,3))
PY
