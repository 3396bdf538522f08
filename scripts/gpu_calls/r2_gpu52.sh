#!/bin/bash
# Round-2 GPU call 52: whole GPU suite + smoke on the final tree (after the LQR batched= flag and the sharding shuffle).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g52_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee -a $O/g52_summary.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/g52_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g52_summary.txt
tail -n 5 $O/g52_pytest.log | tee -a $O/g52_summary.txt
