#!/bin/bash
# Round-2 GPU call 38: new latency-mode defaults (13 warps per SM, pop-size target 15 per SM): queue tests + bench lines + strong leg.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g38_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py tests/test_gpu_parity.py -q -x > $O/g38_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g38_summary.txt
tail -n 4 $O/g38_pytest.log
( time timeout 900 python bench.py ) > $O/g38_bench_default.json 2> $O/g38_bench_default.err; echo "bench default rc=$?" | tee -a $O/g38_summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/g38_bench_driver.json 2> $O/g38_bench_driver.err; echo "bench driver rc=$?" | tee -a $O/g38_summary.txt
python - <<'PY' | tee -a gpurun_out/g38_summary.txt
import json
for f in ("driver","default"):
    d=json.loads(open(f"gpurun_out/g38_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.1f M/s ms/step %.3f lat %.2f ms e2e %.1f frac %.3f traffic x%.2f launches %d parity same %.4f cost %.4f strong %.1f (%.2f ms)" % (
        d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["roofline"]["other"]["traffic_over_algorithmic"],
        d["gpu_launches"], d["parity"]["same_iterations"], d["parity"]["cost_within_1e-4"], d["extra"]["strong"]["value"]/1e6, d["extra"]["strong"]["ms_per_step"]))
    for k,v in d["extra"]["workloads"].items():
        print("   ", k, "value %.4g %s" % (v.get("value",0), v.get("unit")), "frac", round((v.get("roofline") or {}).get("frac") or 0,3), "cpu %.4g" % ((v.get("cpu_baseline") or {}).get("value") or 0))
PY
