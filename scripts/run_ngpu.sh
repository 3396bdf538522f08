#!/bin/bash
# scripts/run_ngpu.sh N tag [bench args]: torchrun bench.py on N GPUs, keep the JSON line in gpurun_out/bench_<tag>.json
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" 2> gpurun_out/bench_$TAG.err | grep "^{" > gpurun_out/bench_$TAG.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("$TAG", d["n_gpus"], "value %.1fM/s" % (d["value"] / 1e6), "%.3f ms/step" % d["ms_per_step"], "e2e %.1fM/s" % (d["e2e"]["value"] / 1e6), "per rank", d.get("per_rank", {}).get("ranks") if d.get("per_rank") else None)
PY
