#!/bin/bash
# Round-2 GPU call 17 (2 GPUs): NCCL test of the sharded solve, bench at N=2 (weak + strong legs), bench at N=1 on the same box.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | tee $O/g17_summary.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/g17_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $O/g17_summary.txt
tail -n 15 $O/g17_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 64 --warmup 3 > $O/g17_bench_2gpu.json 2> $O/g17_bench_2gpu.err; echo "bench2 rc=$?" | tee -a $O/g17_summary.txt
tail -n 3 $O/g17_bench_2gpu.err
timeout 600 python bench.py --no-extra > $O/g17_bench_1gpu.json 2> $O/g17_bench_1gpu.err; echo "bench1 rc=$?" | tee -a $O/g17_summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/g17_bench_ref_2gpu.json 2> $O/g17_bench_ref_2gpu.err; echo "ref2 rc=$?" | tee -a $O/g17_summary.txt
python - <<'PY' | tee -a gpurun_out/g17_summary.txt
import json
for f in ("gpurun_out/g17_bench_1gpu.json","gpurun_out/g17_bench_2gpu.json","gpurun_out/g17_bench_ref_2gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f M/s"%(d["value"]/1e6), "ms/step", round(d["ms_per_step"],3), "e2e %.1f"%(d["e2e"]["value"]/1e6), "seq", d.get("sequential",{}).get("latency_ms_per_batch"), "strong", (d.get("extra") or {}).get("strong",{}).get("value"), (d.get("extra") or {}).get("strong",{}).get("ms_per_step"), "per_rank", d.get("per_rank"))
    except Exception as e: print(f, "FAILED", e)
PY
