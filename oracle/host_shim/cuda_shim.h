/* oracle/host_shim/cuda_shim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Lets the reference engine's CUDA sources (GPU_Rendering_Engine/Source, compiled where they lie under
 * /root/reference) build as plain host C++ with g++: execution-space qualifiers vanish, the built-in
 * index variables become thread-locals that the harness sets while it walks the launch grid with OpenMP,
 * and the two atomics the reference uses map onto GCC __atomic builtins. Force-included with
 * `g++ -x c++ -include cuda_shim.h`. The result is the reference's own algorithm on host cores: it pins
 * the CPU restatement in oracle/rlpt_oracle.cpp and is the `--impl reference` arm of bench.py. */
#ifndef RLPT_ORACLE_CUDA_SHIM_H
#define RLPT_ORACLE_CUDA_SHIM_H
#include <cmath>
#include <math.h>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#define __host__
#define __device__
#define __global__
#define RLPT_REF_HOST 1

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct rlpt_uint3 { unsigned x, y, z; };
extern thread_local rlpt_uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

static inline float max(float a, float b) { return a > b ? a : b; }
static inline float min(float a, float b) { return a < b ? a : b; }

static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, int v) { return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_RELAXED); }
static inline float atomicExch(float* p, float v) {
    uint32_t nv, ov; std::memcpy(&nv, &v, 4);
    ov = __atomic_exchange_n(reinterpret_cast<uint32_t*>(p), nv, __ATOMIC_RELAXED);
    float o; std::memcpy(&o, &ov, 4); return o;
}
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
#endif
