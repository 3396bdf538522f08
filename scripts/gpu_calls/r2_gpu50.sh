#!/bin/bash
# Round-2 GPU call 50: LQR tests + smoke after the batched= flag on LQR.solve.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/g50_summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "lqr or LQR or batched" 2>&1 | tail -2 | tee -a gpurun_out/g50_summary.txt
