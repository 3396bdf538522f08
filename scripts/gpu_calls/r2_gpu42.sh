#!/bin/bash
# Round-2 GPU call 42: new throughput default (pop-size target = SMs / 4) + explicit pipeline mode in bench.py: queue tests, both bench lines.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g42_summary.txt
timeout 600 python -m pytest tests/test_gpu_queue.py -q > $O/g42_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g42_summary.txt
tail -n 3 $O/g42_pytest.log
( time timeout 900 python bench.py ) > $O/g42_bench_default.json 2> $O/g42_bench_default.err; echo "bench default rc=$?" | tee -a $O/g42_summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/g42_bench_driver.json 2> $O/g42_bench_driver.err; echo "bench driver rc=$?" | tee -a $O/g42_summary.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/g42_bench_reference.json 2> $O/g42_bench_reference.err
python - <<'PY' | tee -a gpurun_out/g42_summary.txt
import json
r=json.loads(open("gpurun_out/g42_bench_reference.json").read().strip().splitlines()[-1])
for f in ("driver","default"):
    d=json.loads(open(f"gpurun_out/g42_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.1f M/s ms/step %.3f lat %.2f ms e2e %.1f frac %.3f traffic x%.2f launches %d parity same %.4f cost %.4f strong %.1f (%.2f ms) | ratio %.0f e2e ratio %.0f same_config %s clocks %s" % (
        d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["roofline"]["other"]["traffic_over_algorithmic"],
        d["gpu_launches"], d["parity"]["same_iterations"], d["parity"]["cost_within_1e-4"], d["extra"]["strong"]["value"]/1e6, d["extra"]["strong"]["ms_per_step"], d["value"]/r["value"], d["e2e"]["value"]/r["value"], d["config"]==r["config"], d["clocks"]))
PY
