#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference engine (GPU_Rendering_Engine/Source/*.cu, read from /root/reference where
# it lies; not the reference's CMake build) plus oracle/ref_harness.cu into two checkers under oracle/_ref/:
#   libref_host.so   g++ + oracle/host_shim  (runs on any CPU; used by the CPU tests and bench.py --impl reference)
#   libref_cuda.so   nvcc -rdc=true sm_100a  (the reference's own kernels; GPU box only)
# Excluded on purpose: main.cu and deep_learning/* (DyNet, SDL window), sdl/sdl_screen.cpp (SDL2), voronoi (debug view).
# usage: oracle/build_ref.sh [host|cuda|all] [W H SPP [SUFFIX]]   (resolution and spp are compile-time in the reference)
#        RLPT_ORACLE_ENV=1.0f oracle/build_ref.sh cuda 512 512 32 _env1   (ENVIRONMENT_LIGHT is compile-time too: the Medieval_House preset)
set -euo pipefail
cd "$(dirname "$0")"
WHAT="${1:-all}"; W="${2:-512}"; H="${3:-512}"; SPP="${4:-32}"; SUFFIX="${5:-}"
REF="${RLPT_REFERENCE_ROOT:-/root/reference}"
G="$REF/GPU_Rendering_Engine"; S="$G/Source"
if [ ! -d "$S" ]; then echo "build_ref.sh: reference not present at $REF (nothing to build)"; exit 0; fi
OUT=_ref; HOBJ="$OUT/hobj$SUFFIX"; COBJ="$OUT/cobj$SUFFIX"; mkdir -p "$HOBJ" "$COBJ"
DEFS="-DRLPT_ORACLE_W=$W -DRLPT_ORACLE_H=$H -DRLPT_ORACLE_SPP=$SPP"
if [ -n "${RLPT_ORACLE_ENV:-}" ]; then DEFS="$DEFS -DRLPT_ORACLE_ENV=$RLPT_ORACLE_ENV"; fi
INC="-Iref_overrides -I$G/glm -I$S -I$S/constants -I$S/rays -I$S/objects -I$S/lights -I$S/scenes -I$S/utils -I$S/radiance_volumes -I$S/path_tracing -I$S/sdl"
SRCS="rays/ray camera objects/triangle objects/material objects/surface objects/shape objects/object_importer lights/area_light
      scenes/scene scenes/cornell_box_scene utils/hemisphere_helpers utils/stack utils/printing radiance_volumes/radiance_volume
      radiance_volumes/radiance_tree radiance_volumes/radiance_volume_comparator path_tracing/default_path_tracing
      path_tracing/reinforcement_path_tracing"

build_host() {
  local objs=""
  for f in $SRCS; do
    local ext=cu; [ -f "$S/$f.cu" ] || ext=cpp
    # object_importer.cu and radiance_tree.cu contain functions that fall off the end of a non-void function
    # (object_importer.cu:8-89, radiance_tree.cu:100-115): undefined behaviour above -O0 with gcc 13.
    # gcc 13 also plants a trap at that point even at -O0 unless -fno-unreachable-traps is given.
    local opt=-O2; case "$f" in objects/object_importer|radiance_volumes/radiance_tree|scenes/scene) opt="-O0 -fno-unreachable-traps";; esac
    g++ -std=c++17 $opt -fPIC -fopenmp -w -x c++ -include host_shim/cuda_shim.h -Ihost_shim $DEFS $INC -c "$S/$f.$ext" -o "$HOBJ/$(basename $f).o" &
    objs="$objs $HOBJ/$(basename $f).o"
  done
  g++ -std=c++17 -O2 -fPIC -fopenmp -w -x c++ -include host_shim/cuda_shim.h -Ihost_shim $DEFS $INC -I"$S/radiance_volumes" -c ref_host_radiance_map.cpp -o "$HOBJ/radiance_map.o" &
  g++ -std=c++17 -O2 -fPIC -fopenmp -w -x c++ -include host_shim/cuda_shim.h -Ihost_shim $DEFS $INC -c ref_harness.cu -o "$HOBJ/ref_harness.o" &
  wait
  g++ -shared -fopenmp -o "$OUT/libref_host$SUFFIX.so" $objs "$HOBJ/radiance_map.o" "$HOBJ/ref_harness.o"
  echo "built $OUT/libref_host$SUFFIX.so (${W}x${H}, ${SPP} spp)"
}

build_cuda() {
  local objs=""
  local NV="nvcc -std=c++17 -rdc=true -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fno-unreachable-traps -w $DEFS $INC"
  for f in $SRCS radiance_volumes/radiance_map utils/cuda_helpers; do
    local ext=cu; [ -f "$S/$f.cu" ] || ext=cpp
    $NV -x cu -c "$S/$f.$ext" -o "$COBJ/$(basename $f).o" &
    objs="$objs $COBJ/$(basename $f).o"
  done
  $NV -c ref_harness.cu -o "$COBJ/ref_harness.o" &
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -dlink $objs "$COBJ/ref_harness.o" -o "$COBJ/dlink.o"
  g++ -shared -o "$OUT/libref_cuda$SUFFIX.so" $objs "$COBJ/ref_harness.o" "$COBJ/dlink.o" -L/usr/local/cuda/lib64 -lcudart -lcudadevrt
  echo "built $OUT/libref_cuda$SUFFIX.so (${W}x${H}, ${SPP} spp)"
}

case "$WHAT" in
  host) build_host;;
  cuda) build_cuda;;
  all)  build_host; build_cuda;;
  *) echo "usage: $0 [host|cuda|all] [W H SPP]"; exit 2;;
esac
