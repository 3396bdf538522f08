#!/bin/bash
# Round-2 GPU call 19: latency mode at 14 warps/SM by default -- queue tests, default bench line, launch list of the bench command.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_queue.py tests/test_gpu_parity.py -q -x > $O/g19_pytest.log 2>&1; echo "pytest rc=$?" | tee $O/g19_summary.txt
tail -n 6 $O/g19_pytest.log
( time timeout 900 python bench.py ) > $O/g19_bench_default.json 2> $O/g19_bench_default.err; echo "bench rc=$?" | tee -a $O/g19_summary.txt
tail -n 4 $O/g19_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/g19_bench_reference.json 2> $O/g19_bench_reference.err; echo "ref rc=$?" | tee -a $O/g19_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/g19_launches_bench.csv python bench.py --steps 8 --warmup 3 --no-extra --no-cpu-baseline --no-clock-sampler > $O/g19_bench_under_ncu.log 2>&1; echo "ncu rc=$?" | tee -a $O/g19_summary.txt
python - <<'PY' | tee -a gpurun_out/g19_summary.txt
import json
d=json.loads(open("gpurun_out/g19_bench_default.json").read().strip().splitlines()[-1])
print("value %.1f M/s ms/step %.3f lat %.2f ms e2e %.1f frac %.3f traffic %.2f GB parity %s strong %.1f launches %d" % (d["value"]/1e6, d["ms_per_step"], d["sequential"]["latency_ms_per_batch"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["roofline"]["traffic"]/1e9, {k: round(v,4) for k,v in d["parity"].items() if isinstance(v,float)}, d["extra"]["strong"]["value"]/1e6, d["gpu_launches"]))
for k,v in d["extra"]["workloads"].items():
    print(k, "value %.4g %s" % (v.get("value",0), v.get("unit")), "frac", round((v.get("roofline") or {}).get("frac") or 0,3), "traffic", (v.get("roofline") or {}).get("traffic"), "cpu", (v.get("cpu_baseline") or {}).get("value"), "parity", {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in (v.get("parity") or {}).items() if kk not in ("against","max_rel_error")})
r=json.loads(open("gpurun_out/g19_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm value %.3f M/s cores %d same_config %s" % (r["value"]/1e6, r["cpu_baseline"]["cores"], r["config"]==d["config"]))
PY
