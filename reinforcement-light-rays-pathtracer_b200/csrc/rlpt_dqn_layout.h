// rlpt_dqn_layout.h -- where a weight of the Neural-Q network lives in the packed bf16 operands k_dqn_forward / k_dqn_backward stream (plain C++: included by
// rlpt_dqn.h for the kernels, and by tests/dqn_layout_host.cpp, which checks on the host that the layout is a gap-free bijection made of contiguous chunks).
#pragma once
#include <stddef.h>
#if defined(__CUDACC__)
#define RLPT_LAYOUT_HD __host__ __device__
#else
#define RLPT_LAYOUT_HD
#endif
namespace rlpt {
#ifndef RLPT_DQN_KC_A
#define RLPT_DQN_KC_A 112        // inputs per weight chunk of the 208-input layers (2 and 4; 208 = 112 + 96)
#endif
#ifndef RLPT_DQN_KC_B
#define RLPT_DQN_KC_B 80         // ... of the 304-input layer (3; 304 = 3 x 80 + 64)
#endif
constexpr int DQ_LAYOUT_K3 = 304;                                          // (= DQ_K3 of rlpt_dqn.h)
// Inputs per weight chunk. One cp.async.bulk occupies the SM's copy engine for >= ~360 cycles whatever its size, and copies are served one after the other
// (scratch/ubench/copy_bw.cu: 5 KB copies 14 B/cycle, 20 KB 57, 33 KB 91, >= 40 KB 113 B/cycle per SM): few large chunks, not many small ones.
RLPT_LAYOUT_HD constexpr int dq_kc(int k_pad) { return k_pad == DQ_LAYOUT_K3 ? RLPT_DQN_KC_B : RLPT_DQN_KC_A; }
static_assert(RLPT_DQN_KC_A % 16 == 0 && RLPT_DQN_KC_B % 16 == 0, "chunks are whole MMA K steps");
constexpr int DQ_L2_SPLIT = 160;                                           // layer 2's first N part
// Byte offset of weight (row = output unit, k = input unit) in a packed operand of n_pad x k_pad bf16. The rows are cut into N parts (n_split > 0: rows
// [0, n_split) and [n_split, n_pad); 0: one part), each part into K chunks of dq_kc(k_pad) inputs (the last one shorter); a chunk is one contiguous block
// = one bulk copy = the B operand of kw / 16 MMAs. Inside a chunk: canonical K-major no-swizzle form, 8x8 core matrices of 128 bytes, K-adjacent ones 128 bytes
// apart (LBO), 8-row groups kw * 16 bytes apart (SBO).
RLPT_LAYOUT_HD inline size_t wpack_offset(int n_split, int row, int k, int n_pad, int k_pad) {
    const int n0 = (n_split > 0 && row >= n_split) ? n_split : 0, rows = n_split > 0 ? (row >= n_split ? n_pad - n_split : n_split) : n_pad;
    const int KC = dq_kc(k_pad), kc = k / KC, kw = (k_pad - kc * KC) < KC ? (k_pad - kc * KC) : KC, r = row - n0, kk = k - kc * KC;
    return (size_t)n0 * k_pad * 2 + (size_t)rows * KC * 2 * kc + (size_t)(r >> 3) * ((size_t)kw * 16) + (size_t)(kk >> 3) * 128 + (size_t)(r & 7) * 16 + (size_t)(kk & 7) * 2;
}
}  // namespace rlpt
