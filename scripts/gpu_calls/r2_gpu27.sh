#!/bin/bash
# Round-2 GPU call 27 (8 GPUs): end-to-end path with each rank bound to the NUMA node of its GPU, against unbound.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g27_summary.txt
nvidia-smi topo -m > $O/g27_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name" >> $O/g27_topo.txt
for mode in numa nonuma; do
  if [ $mode = nonuma ]; then export TFMPC_BENCH_NO_NUMA=1; else unset TFMPC_BENCH_NO_NUMA; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 8 --steps 64 --warmup 3 --no-strong > $O/g27_bench_8gpu_$mode.json 2> $O/g27_bench_8gpu_$mode.err; echo "$mode rc=$?" | tee -a $O/g27_summary.txt
done
unset TFMPC_BENCH_NO_NUMA
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus 4 --steps 64 --warmup 3 --no-strong > $O/g27_bench_4gpu_numa.json 2> $O/g27_bench_4gpu_numa.err; echo "4gpu rc=$?" | tee -a $O/g27_summary.txt
python - <<'PY' | tee -a gpurun_out/g27_summary.txt
import json
for f in ("8gpu_numa","8gpu_nonuma","4gpu_numa"):
    try:
        d=json.loads(open(f"gpurun_out/g27_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1f M/s e2e %.1f M/s" % (d["value"]/1e6, d["e2e"]["value"]/1e6), d["e2e"].get("numa"))
    except Exception as e: print(f, "FAILED", e)
PY
