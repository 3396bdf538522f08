#!/bin/bash
# Round-2 GPU call 49 (4 GPUs): final tree at N = 4, the driver's invocation.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29931 bench.py --gpus 4 --steps 20 --warmup 5 > $O/g49_bench_4gpu.json 2> $O/g49_bench_4gpu.err; echo "N=4 rc=$?" | tee $O/g49_summary.txt
python - <<'PY' | tee -a gpurun_out/g49_summary.txt
import json
d=json.loads(open("gpurun_out/g49_bench_4gpu.json").read().strip().splitlines()[-1])
print("N=4 value %.1f M/s ms/step %.3f e2e %.1f lat %.2f strong %.1f (%.2f ms) gather: %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["sequential"]["latency_ms_per_batch"], d["extra"]["strong"]["value"]/1e6, d["extra"]["strong"]["ms_per_step"], d["details"]["parallelism"][-110:]))
PY
