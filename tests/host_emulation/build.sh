#!/bin/bash
# Builds the host-emulation harness (TEST INFRASTRUCTURE: runs the __host__ __device__ per-problem code of
# tfmpc_b200/csrc/small_core.cuh and the warp-level queue solver of queue_core.cuh on the CPU so the GPU-less CI can
# exercise the device logic).
set -e
cd "$(dirname "$0")"
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
CUDA_INC=/usr/local/cuda/include
SRC=../../tfmpc_b200/csrc
for p in f32 f64; do
  D=""; [ "$p" = f64 ] && D="-DTFMPC_F64"
  out=libemul_$p.so
  if [ ! -f "$out" ] || [ emul.cpp -nt "$out" ] || [ emul_queue.cpp -nt "$out" ] || [ $SRC/small_core.cuh -nt "$out" ] || [ $SRC/common.cuh -nt "$out" ] \
     || [ $SRC/queue_core.cuh -nt "$out" ] || [ $SRC/warp_rt.cuh -nt "$out" ]; then
    "$CXX" -O2 -std=c++17 -fPIC -shared -pthread -ffp-contract=off -Wno-unknown-pragmas $D -I"$CUDA_INC" emul.cpp emul_queue.cpp -o "$out"
  fi
done
