#!/bin/bash
# Round-2 GPU call 39 (2 GPUs): final tree under torchrun -- NCCL test, driver-style invocations at N = 1 and N = 2, reference arm at N = 2.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g39_summary.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -n 2 | tee -a $O/g39_summary.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra > $O/g39_bench_1gpu.json 2> $O/g39_bench_1gpu.err; echo "N=1 rc=$?" | tee -a $O/g39_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > $O/g39_bench_2gpu.json 2> $O/g39_bench_2gpu.err; echo "N=2 rc=$?" | tee -a $O/g39_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29912 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/g39_bench_ref_2gpu.json 2> $O/g39_bench_ref_2gpu.err; echo "ref N=2 rc=$?" | tee -a $O/g39_summary.txt
python - <<'PY' | tee -a gpurun_out/g39_summary.txt
import json
b=None
for n in (1,2):
    d=json.loads(open(f"gpurun_out/g39_bench_{n}gpu.json").read().strip().splitlines()[-1])
    if n==1: b=d
    print("N=%d value %.1f M/s (eff %.3f) e2e %.1f (eff %.3f) lat %.2f strong %.1f (%.2f ms)" % (n, d["value"]/1e6, d["value"]/(n*b["value"]), d["e2e"]["value"]/1e6, d["e2e"]["value"]/(n*b["e2e"]["value"]), d["sequential"]["latency_ms_per_batch"], d["extra"]["strong"]["value"]/1e6, d["extra"]["strong"]["ms_per_step"]))
r=json.loads(open("gpurun_out/g39_bench_ref_2gpu.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/g39_bench_2gpu.json").read().strip().splitlines()[-1])
print("reference arm N=2: %.3f M/s same_config %s" % (r["value"]/1e6, r["config"]==d["config"]))
PY
