#!/bin/bash
# Round-2 GPU call 32: final tree -- build check, smoke, whole GPU suite, the driver's bench invocation and the default one, reference arm.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g32_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee -a $O/g32_summary.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > $O/g32_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g32_summary.txt
tail -n 5 $O/g32_pytest.log | tee -a $O/g32_summary.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/g32_bench_reference_driver.json 2> $O/g32_bench_reference_driver.err; echo "ref rc=$?" | tee -a $O/g32_summary.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/g32_bench_driver.json 2> $O/g32_bench_driver.err; echo "bench driver rc=$?" | tee -a $O/g32_summary.txt
tail -n 4 $O/g32_bench_driver.err | tee -a $O/g32_summary.txt
( time timeout 900 python bench.py ) > $O/g32_bench_default.json 2> $O/g32_bench_default.err; echo "bench default rc=$?" | tee -a $O/g32_summary.txt
python - <<'PY' | tee -a gpurun_out/g32_summary.txt
import json
r=json.loads(open("gpurun_out/g32_bench_reference_driver.json").read().strip().splitlines()[-1])
for f in ("driver","default"):
    d=json.loads(open(f"gpurun_out/g32_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.1f M/s ms/step %.3f lat %.2f ms e2e %.1f frac %.3f traffic x%.2f launches %d parity same %.4f cost %.4f strong %.1f | ratio vs ref %.1f e2e ratio %.1f same_config %s" % (
        d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["roofline"]["other"]["traffic_over_algorithmic"],
        d["gpu_launches"], d["parity"]["same_iterations"], d["parity"]["cost_within_1e-4"], d["extra"]["strong"]["value"]/1e6, d["value"]/r["value"], d["e2e"]["value"]/r["value"], d["config"]==r["config"]))
    for k,v in d["extra"]["workloads"].items():
        print("   ", k, "value %.4g %s" % (v.get("value",0), v.get("unit")), "frac", round((v.get("roofline") or {}).get("frac") or 0,3), "cpu %.4g" % ((v.get("cpu_baseline") or {}).get("value") or 0))
PY
