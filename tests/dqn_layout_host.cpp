// tests/dqn_layout_host.cpp -- host check of the packed weight layout the Neural-Q kernels stream (csrc/rlpt_dqn_layout.h, compiled here by g++ as it is):
// for an operand of n_pad x k_pad bf16 with N split n_split, wpack_offset must be (1) a bijection onto the even bytes of [0, 2 n_pad k_pad), (2) made of
// contiguous chunks -- part p, K chunk c occupies exactly [base, base + rows * kw * 2) with the chunks of a part, and the parts, one behind the other (that is
// what lets one cp.async.bulk bring a chunk and k_dqn_backward / k_dqn_forward address it by a base pointer plus rows * k0), and (3) inside a chunk the
// canonical K-major no-swizzle tcgen05 operand form with SBO = kw * 16 bytes, LBO = 128. Returns the number of violations.
#include <stdint.h>
#include <vector>
#include "rlpt_dqn_layout.h"
extern "C" long dqn_layout_check(int n_split, int n_pad, int k_pad) {
    using namespace rlpt;
    long bad = 0;
    std::vector<uint8_t> seen((size_t)n_pad * k_pad, 0);
    const int KC = dq_kc(k_pad);
    for (int row = 0; row < n_pad; ++row) for (int k = 0; k < k_pad; ++k) {
        const size_t off = wpack_offset(n_split, row, k, n_pad, k_pad);
        if (off % 2 || off / 2 >= seen.size() || seen[off / 2]++) { ++bad; continue; }
        const int part = (n_split > 0 && row >= n_split) ? 1 : 0, n0 = part ? n_split : 0, rows = n_split > 0 ? (part ? n_pad - n_split : n_split) : n_pad;
        const int c = k / KC, k0 = c * KC, kw = k_pad - k0 < KC ? k_pad - k0 : KC, r = row - n0, kk = k - k0;
        const size_t base = (size_t)n0 * k_pad * 2 + (size_t)rows * k0 * 2;                       // chunks of a part, and the parts, one behind the other
        if (off < base || off >= base + (size_t)rows * kw * 2) ++bad;                               // inside its chunk
        if (off - base != (size_t)(r >> 3) * kw * 16 + (size_t)(kk >> 3) * 128 + (size_t)(r & 7) * 16 + (size_t)(kk & 7) * 2) ++bad;
        if (kw % 16 || rows % 8) ++bad;                                                            // whole MMA K steps, whole 8-row groups
    }
    for (uint8_t v : seen) if (v != 1) ++bad;
    return bad;
}
extern "C" int dqn_layout_kc(int k_pad) { return rlpt::dq_kc(k_pad); }
extern "C" int dqn_layout_split() { return rlpt::DQ_L2_SPLIT; }
