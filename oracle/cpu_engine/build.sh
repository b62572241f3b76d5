#!/usr/bin/env bash
# oracle/cpu_engine/build.sh -- TEST / BASELINE INFRASTRUCTURE ONLY.
# Builds the reference's Old_CPU_Rendering_Engine (BASELINE.json configs[0]) from where it lies under /root/reference -- its own
# CMake build is not used (icc, SDL2 and libiomp5 are absent) -- plus cpu_engine_harness.cpp into baseline/_ref/ (git-ignored; it
# travels to the GPU box like any built .so): libcpu_engine_b2.so (MAX_RAY_BOUNCES 2 as committed) and libcpu_engine_b80.so
# (80, the GPU engine's setting). g++ -O3 -fopenmp stands in for icc -O3 -qopenmp -no-prec-div (CMakeLists.txt:5-8).
# usage: oracle/cpu_engine/build.sh [SPP]
set -euo pipefail
cd "$(dirname "$0")"
REF="${RLPT_REFERENCE_ROOT:-/root/reference}"; C="$REF/Old_CPU_Rendering_Engine"; S="$C/Source"
if [ ! -d "$S" ]; then echo "cpu_engine/build.sh: reference not present at $REF (nothing to build)"; exit 0; fi
OUT=../../baseline/_ref; mkdir -p "$OUT"
SPP="${1:-16}"
INC="-Ioverrides -I$C/glm -I$S -I$S/constants -I$S/objects -I$S/scenes -I$S/utils -I$S/lights -I$S/sdl -I$S/rays -I$S/radiance_volumes -I$S/path_tracing"
SRCS="rays/ray camera objects/material objects/shape objects/triangle objects/surface lights/area_light lights/area_light_plane
      scenes/cornell_box_scene utils/hemisphere_helpers utils/printing path_tracing/default_path_tracing"
for B in 2 80; do
  OBJ="$OUT/cpu_obj_b$B"; mkdir -p "$OBJ"; objs=""
  for f in $SRCS; do
    g++ -std=c++11 -O3 -fopenmp -fPIC -w -include cpu_engine_shim.h -DRLPT_CE_BOUNCES=$B -DRLPT_CE_SPP=$SPP $INC -c "$S/$f.cpp" -o "$OBJ/$(basename $f).o" &
    objs="$objs $OBJ/$(basename $f).o"
  done
  g++ -std=c++11 -O3 -fopenmp -fPIC -w -include cpu_engine_shim.h -DRLPT_CE_BOUNCES=$B -DRLPT_CE_SPP=$SPP $INC -c cpu_engine_harness.cpp -o "$OBJ/harness.o" &
  wait
  g++ -shared -fopenmp -o "$OUT/libcpu_engine_b$B.so" $objs "$OBJ/harness.o"
  echo "built baseline/_ref/libcpu_engine_b$B.so (512x512, $SPP spp, $B bounces)"
done
