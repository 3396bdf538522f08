#!/bin/bash
# Round-2 GPU call 10: whole GPU suite on the current tree + phase-resolved scheduling trace of a lone C3 batch.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/g10_pytest.log 2>&1; echo "pytest rc=$?" | tee $O/g10_summary.txt
tail -n 15 $O/g10_pytest.log
TFMPC_QUEUE_WTARGET=148 timeout 300 python scripts/queue_trace.py --tag g10_wt148 --streams 8 --rounds 2 2>&1 | tail -n 2 | tee -a $O/g10_summary.txt
TFMPC_QUEUE_WTARGET=2664 timeout 300 python scripts/queue_trace.py --tag g10_wt2664 --streams 8 --rounds 2 2>&1 | tail -n 2 | tee -a $O/g10_summary.txt
