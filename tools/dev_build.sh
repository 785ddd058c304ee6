#!/bin/bash
# Development build: nvcc --split-compile (parallel back end, ~2 min instead of ~5-8).  Register allocation differs
# slightly from the release build of __graft_entry__.build(), so numbers for profiles/ are taken with the release build.
cd "$(dirname "$0")/.."
exec nvcc -std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -shared \
    --split-compile 0 -I include -o deeplearning_b200/libdlb200.so deeplearning_b200/csrc/dlb200.cu "$@"
