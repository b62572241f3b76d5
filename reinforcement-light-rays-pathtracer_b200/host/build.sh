#!/usr/bin/env bash
# Builds the C++ host mirror: librlpt_host.so (Scene/Camera/SDLScreen/Renderer/RadianceMap over the C ABI + the NCCL
# all-reduce hook) and the rlpt_example program. -ffp-contract=off: scene arithmetic rounds like the reference's host code.
set -euo pipefail
cd "$(dirname "$0")"
LIB=../lib; mkdir -p "$LIB" ../build
CUDA="${CUDA_HOME:-/usr/local/cuda}"
CXXFLAGS="-std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -Wall -I$CUDA/include"
g++ $CXXFLAGS -c rlpt_host.cpp -o ../build/rlpt_host.o
NCCL_OBJ=""; NCCL_LIB=""
if [ -f /usr/include/nccl.h ] && ls /usr/lib/x86_64-linux-gnu/libnccl.so* >/dev/null 2>&1; then
  g++ $CXXFLAGS -c nccl_hook.cpp -o ../build/nccl_hook.o; NCCL_OBJ=../build/nccl_hook.o; NCCL_LIB="-lnccl"
fi
g++ -shared -o "$LIB/librlpt_host.so" ../build/rlpt_host.o $NCCL_OBJ -L"$LIB" -lrlpt -L"$CUDA/lib64" -lcudart $NCCL_LIB -Wl,-rpath,'$ORIGIN'
if [ -n "$NCCL_LIB" ]; then
  g++ $CXXFLAGS example_main.cpp -o "$LIB/rlpt_example" -L"$LIB" -lrlpt_host -lrlpt -L"$CUDA/lib64" -lcudart $NCCL_LIB -lpthread -Wl,-rpath,'$ORIGIN'
fi
echo "built $(cd "$LIB" && pwd)/librlpt_host.so"
