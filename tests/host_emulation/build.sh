#!/bin/bash
# Builds the host-emulation harness (TEST INFRASTRUCTURE: runs the __host__ __device__ per-problem code of
# tfmpc_b200/csrc/small_core.cuh on the CPU so the GPU-less CI can exercise the device logic).
set -e
cd "$(dirname "$0")"
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
CUDA_INC=/usr/local/cuda/include
for p in f32 f64; do
  D=""; [ "$p" = f64 ] && D="-DTFMPC_F64"
  out=libemul_$p.so
  if [ ! -f "$out" ] || [ emul.cpp -nt "$out" ] || [ ../../tfmpc_b200/csrc/small_core.cuh -nt "$out" ] || [ ../../tfmpc_b200/csrc/common.cuh -nt "$out" ]; then
    "$CXX" -O2 -std=c++17 -fPIC -shared -ffp-contract=off -Wno-unknown-pragmas $D -I"$CUDA_INC" emul.cpp -o "$out"
  fi
done
