#!/usr/bin/env bash
# Builds librlpt.so (the C-ABI library of include/rlpt.h) for sm_100a, in-tree.
#   -lineinfo            so ncu's source page maps to these files
#   -Xcompiler -ffp-contract=off   host code: one rounding per operator (radiance-map construction must match the
#                        reference's host arithmetic bit for bit); device code spells its roundings out with intrinsics
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p lib build
NVCC="${NVCC:-nvcc}"
FLAGS="${RLPT_EXTRA_NVCC:-} -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -Xptxas -v -ccbin g++"
pids=()
for f in rlpt_kernels rlpt_bvh rlpt_capi rlpt_dqn; do
  $NVCC $FLAGS -c csrc/$f.cu -o build/$f.o 2> build/$f.ptxas.log & pids+=($!)
done
g++ -std=c++17 -O2 -fPIC -fopenmp -ffp-contract=off -fno-fast-math -c csrc/rlpt_radiance_host.cpp -o build/rlpt_radiance_host.o & pids+=($!)
for p in "${pids[@]}"; do wait $p || { cat build/*.ptxas.log | grep -iE "error|fatal" -A3 | head -60; exit 1; }; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -ccbin g++ -o lib/${RLPT_LIB_NAME:-librlpt.so} build/rlpt_kernels.o build/rlpt_bvh.o build/rlpt_capi.o build/rlpt_dqn.o build/rlpt_radiance_host.o -lcudart -lgomp
echo "built $(pwd)/lib/librlpt.so"
