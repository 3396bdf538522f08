#!/bin/bash
# Round-2 GPU call 51 (2 GPUs): the NCCL sharding test with the balancing permutation.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15 | tee gpurun_out/g51_summary.txt
