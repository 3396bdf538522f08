#!/bin/bash
# Round-2 GPU call 28 (8 GPUs): the box's aggregate pinned-copy bandwidth with the end-to-end path's buffer sizes, N = 1, 2, 4, 8.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g28_pcie.jsonl
python scripts/pcie_aggregate.py >> $O/g28_pcie.jsonl 2>/dev/null
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900 + n)) scripts/pcie_aggregate.py >> $O/g28_pcie.jsonl 2>/dev/null
done
cat $O/g28_pcie.jsonl
