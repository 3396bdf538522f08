#!/bin/bash
# Round-2 GPU call 15: the whole GPU suite (new: full-size parity, fast-math A/B, retry fixtures, device plant noise, start RNG).
mkdir -p gpurun_out
O=gpurun_out
nproc
( time timeout 1700 python -m pytest tests -m gpu -q --durations=12 ) > $O/g15_pytest.log 2>&1; echo "pytest rc=$?" | tee $O/g15_summary.txt
tail -n 40 $O/g15_pytest.log
