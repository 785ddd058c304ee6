#!/bin/sh
# Builds the development-time CPU logic checker (see cuda_emu.h). Not part of the product build.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/_build"
g++ -std=c++20 -O2 -g -fPIC -shared -fvisibility=hidden -DDLB_CUDA_EMU -include "$HERE/cuda_emu.h" -x c++ \
    -I"$HERE/../../include" -o "$HERE/_build/libdlb200_emu.so" "$HERE/../../deeplearning_b200/csrc/dlb200.cu" "$HERE/mnet_stub.cpp" -lpthread
echo "built $HERE/_build/libdlb200_emu.so"
