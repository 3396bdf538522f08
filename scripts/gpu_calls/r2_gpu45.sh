#!/bin/bash
# Round-2 GPU call 45: last validation of the final tree: smoke, whole GPU suite, reference arm + driver invocation.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g45_summary.txt
git -C . log --oneline 2>/dev/null | head -n 1 | tee -a $O/g45_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee -a $O/g45_summary.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > $O/g45_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g45_summary.txt
tail -n 4 $O/g45_pytest.log | tee -a $O/g45_summary.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/g45_bench_reference.json 2> $O/g45_bench_reference.err; echo "ref rc=$?" | tee -a $O/g45_summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/g45_bench_driver.json 2> $O/g45_bench_driver.err; echo "bench rc=$?" | tee -a $O/g45_summary.txt
python - <<'PY' | tee -a gpurun_out/g45_summary.txt
import json
r=json.loads(open("gpurun_out/g45_bench_reference.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/g45_bench_driver.json").read().strip().splitlines()[-1])
print("driver value %.1f M/s ms/step %.3f lat %.2f ms e2e %.1f frac %.3f launches %d parity same %.4f strong %.1f | ratio %.0f e2e ratio %.0f same_config %s" % (
    d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["gpu_launches"], d["parity"]["same_iterations"], d["extra"]["strong"]["value"]/1e6, d["value"]/r["value"], d["e2e"]["value"]/r["value"], d["config"]==r["config"]))
PY
