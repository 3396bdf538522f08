#!/bin/bash
# Round-2 GPU call 24 (8 GPUs): the driver's scaling sequence N = 1, 2, 4, 8 back to back on one box, plus the reference arm at N = 8.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -n 8 > $O/g24_summary.txt
nproc >> $O/g24_summary.txt
timeout 600 python bench.py --no-extra > $O/g24_bench_1gpu.json 2> $O/g24_bench_1gpu.err; echo "N=1 rc=$?" | tee -a $O/g24_summary.txt
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 64 --warmup 3 > $O/g24_bench_${n}gpu.json 2> $O/g24_bench_${n}gpu.err; echo "N=$n rc=$?" | tee -a $O/g24_summary.txt
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/g24_bench_ref_8gpu.json 2> $O/g24_bench_ref_8gpu.err; echo "ref8 rc=$?" | tee -a $O/g24_summary.txt
python - <<'PY' | tee -a gpurun_out/g24_summary.txt
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/g24_bench_{n}gpu.json").read().strip().splitlines()[-1])
        if n==1: base=d
        st=(d.get("extra") or {}).get("strong",{})
        print("N=%d value %.1f M/s (eff %.3f) e2e %.1f (eff %.3f) seq lat %.2f ms strong %.1f M/s (%.2f ms) gather %s" % (n, d["value"]/1e6, d["value"]/(n*base["value"]), d["e2e"]["value"]/1e6, d["e2e"]["value"]/(n*base["e2e"]["value"]), d["sequential"]["latency_ms_per_batch"], st.get("value",0)/1e6, st.get("ms_per_step",0), d["details"]["parallelism"][-90:]))
    except Exception as e: print(n, "FAILED", e)
PY
