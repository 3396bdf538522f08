#!/bin/bash
mkdir -p gpurun_out
export TFMPC_QUEUE_WTARGET=74 TFMPC_QUEUE_PATIENCE=0
TFMPC_B200_LIBDIR=$PWD/ab/gs13 timeout 300 python scripts/queue_trace_all.py --tag g8_gs13 --streams 8 --rounds 4 2>&1 | tail -2
TFMPC_B200_LIBDIR=$PWD/ab/wps16 timeout 300 python scripts/queue_trace_all.py --tag g8_wps16 --streams 8 --rounds 4 2>&1 | tail -2
