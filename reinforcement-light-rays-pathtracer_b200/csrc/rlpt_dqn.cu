// rlpt_dqn.cu -- Neural-Q network forward pass on tcgen05 tensor cores (sm_100a), parameter handling, DyNet text IO.
//
// Replaces: convert_vertices_to_point_coord_system (G/deep_learning/nn_rendering_helpers.cu:280-298, which materialises
// a K-float input per ray: 358 MB at 512^2 x 342) + DQNetwork::network_inference through DyNet
// (N/dq_network.cu:37-49, called at G/deep_learning/neural_q_pathtracer.cu:321-325,440-444,494-499).
//
// One CTA = one tile of 128 rays = the M dimension of tcgen05.mma (cta_group::1), accumulators in TMEM (128 lanes x
// 512 fp32 columns: layer 2 in columns [0,304), layer 3 in [304,512), layer 4 reuses [0,144)). 576 threads per CTA:
// 16 epilogue / layer-1 warps, one MMA warp, one copy warp. Per tile:
//   layer 1   fp32 on the CUDA cores, c1 - M1 x (rlpt_dqn.h), ReLU, written as the bf16 A operand into shared memory
//   layer 2-4 the packed weights (wpack_offset: N parts x K chunks of 27-36 KB) stream from L2 through a two-stage ring of
//             cp.async.bulk copies issued by the copy lane from a chunk table in shared memory; the MMA lane issues kw / 16
//             tcgen05.mma per chunk with the part's full N; tcgen05.commit signals mbarriers (stage free / N part done);
//             the epilogue warps read TMEM with tcgen05.ld.32x32b.x32 (thread t <-> lane t & 127 <-> ray, the four warpgroups
//             take 32-column blocks in turn), add the bias, apply ReLU and write the next layer's A operand (bf16) straight
//             back into shared memory -- activations never touch HBM (the training instantiation also keeps them, and relu'
//             as bit words, for the backward pass)
//   output    Q values, fp32, action-major [144][n] so that producer and consumers are coalesced
// k_dqn_backward (the TD step's backward data path) has the same shape; k_gemm_bf16_tn (weight gradients) is a single-shot
// cp.async GEMM. DESIGN.md section 3b has the measurements that led here.
// Operand layout (both A and B): K-major, SWIZZLE_NONE canonical form: 8x8 core matrices of 128 contiguous bytes
// (8 rows x 16 bytes), K-adjacent core matrices 128 bytes apart (LBO), 8-row groups K_pad*16 bytes apart (SBO; kw*16 inside a
// streamed weight chunk).
#include "rlpt_dqn.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rlpt {

#define DQ_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

// Programmatic dependent launch (the optimiser step is a chain of small kernels): a kernel launched with the programmatic-serialisation attribute
// may become resident while its predecessor still runs; it waits here before it touches anything the predecessor produces, and only then lets
// ITS successor in (so at most two kernels of the chain hold SM resources at a time). Without the attribute both instructions fall through.
#define RLPT_PDL_SYNC() do { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); } while (0)
// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// (shared-memory matrix descriptors, cute::UMMA::SmemDescriptor: start address >> 4 at bit 0, LBO >> 4 at bit 16, SBO >> 4 at bit 32, version 1 at bit 46, no swizzle --
// built as (low, high) words where they are used: umma_bf16_lohi)
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// byte offset of element (row, k) in a K-major canonical operand with K_pad columns
__host__ __device__ __forceinline__ size_t operand_offset(int row, int k, int k_pad) {
    return (size_t)(row >> 3) * ((size_t)k_pad * 16) + (size_t)(k >> 3) * 128 + (size_t)(row & 7) * 16 + (size_t)(k & 7) * 2;
}

// Packed weights of a layer as k_dqn_forward streams them (B operands): the layer's output rows are cut into N parts (layer 2: rows [0,160) and
// [160,304), so that the epilogue of the first part runs under the MMAs of the second; layers 3 and 4: one part), each part into K chunks of dq_kc
// inputs; a chunk is one contiguous block = one bulk copy = the B operand of kw / 16 MMAs with N = the part's rows (large N: an SS-mode MMA re-reads
// its 4 KB of A from shared memory, so N = 64 pieces are shared-memory-bound and issue-bound). Inside a chunk: canonical K-major no-swizzle form,
// 8x8 core matrices of 128 bytes, K-adjacent ones 128 bytes apart (LBO), 8-row groups kw * 16 bytes apart (SBO).
#ifndef RLPT_DQN_STAGES
#define RLPT_DQN_STAGES 2
#endif
#ifndef RLPT_DQN_PIECES
#define RLPT_DQN_PIECES 1
#endif
// (chunk sizes, dq_kc, DQ_L2_SPLIT and wpack_offset: rlpt_dqn_layout.h -- plain C++, also compiled into a host test)
static_assert(DQ_LAYOUT_K3 == DQ_K3, "rlpt_dqn_layout.h tells the 304-input layer by its width");
// The packed weights exist DQ_REPLICAS times in global memory (replica r at byte offset r * stride): CTA b streams replica b % DQ_REPLICAS. All CTAs of a
// full-frame forward walk the chunk stream in step, so without replicas 148 SMs ask the same 128-byte lines of the same L2 slices at the same moment
// (an A/B switch: it made no difference).
#ifndef RLPT_DQN_REPLICAS
#define RLPT_DQN_REPLICAS 1          // measured with 8: 186.3 vs 186.6 us per full-frame forward -- L2 line contention is not what paces the weight stream
#endif
constexpr int DQ_REPLICAS = RLPT_DQN_REPLICAS;
constexpr size_t DQ_W2P_STRIDE = 2 * (size_t)DQ_N2 * DQ_K2 + 128 * 3, DQ_W3P_STRIDE = 2 * (size_t)DQ_N3 * DQ_K3 + 128 * 5, DQ_W4P_STRIDE = 2 * (size_t)DQ_N4 * DQ_K4 + 128 * 7;       // bytes; odd multiples of a line apart

__constant__ float c_dq_cos[DQ_OUT];                                     // cos(theta) of the 144 grid cells (the tracer's table; dqn_upload_cell_cos)
void dqn_upload_cell_cos(const float* cos144) { cudaMemcpyToSymbol(c_dq_cos, cos144, sizeof(float) * DQ_OUT); }
// ------------------------------------------------------------------------------------------------ forward kernel
constexpr int DQ_STAGES = RLPT_DQN_STAGES;
constexpr uint32_t dq_max3(uint32_t a, uint32_t b, uint32_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
constexpr uint32_t DQ_STAGE_BYTES = (dq_max3(DQ_L2_SPLIT * dq_kc(DQ_K2) * 2, DQ_N3 * dq_kc(DQ_K3) * 2, DQ_N4 * dq_kc(DQ_K4) * 2) + 127u) & ~127u;       // largest chunk (160 rows x 112 inputs = 35 KB)
constexpr uint32_t SM_A1 = 0;                                             // 128 x 208 bf16: layer-2 A, later layer-4 A
constexpr uint32_t SM_A2 = SM_A1 + DQ_TILE * DQ_K2 * 2;                   // 128 x 304 bf16: layer-3 A
constexpr uint32_t SM_B0 = SM_A2 + DQ_TILE * DQ_K3 * 2;                   // DQ_STAGES weight chunks
constexpr uint32_t SM_C1 = SM_B0 + DQ_STAGES * DQ_STAGE_BYTES;            // fp32 constants
constexpr uint32_t SM_M1 = SM_C1 + DQ_K2 * 4;
constexpr uint32_t SM_BIAS2 = SM_M1 + DQ_K2 * 3 * 4;
constexpr uint32_t SM_BIAS3 = SM_BIAS2 + DQ_N2 * 4;
constexpr uint32_t SM_BIAS4 = SM_BIAS3 + DQ_N3 * 4;
constexpr uint32_t SM_BAR = SM_BIAS4 + DQ_N4 * 4;                         // 8 mbarriers + the TMEM base address
constexpr uint32_t SM_TABLE = SM_BAR + 160;                                // chunk table (17 x 40 bytes)
constexpr uint32_t SM_TOTAL = SM_TABLE + 17 * 40;
static_assert(SM_TOTAL <= 227 * 1024, "shared-memory budget");
static_assert(SM_BAR % 8 == 0 && SM_B0 % 128 == 0 && DQ_STAGE_BYTES % 128 == 0, "alignment");

#ifndef RLPT_DQN_STAGGER
#define RLPT_DQN_STAGGER 0         // cycles of start-up delay per CTA (blockIdx % 8): tried 3000 -- 215 -> 236 us per full-frame forward
#endif
#ifndef RLPT_DQN_EPI_GROUPS
#define RLPT_DQN_EPI_GROUPS 4
#endif
// Warp roles (round 2): DQ_EPI_GROUPS warpgroups of epilogue / layer-1 threads (thread t works on ray t & 127; TMEM lane quarter = warp & 3, the
// warpgroups take 32-column blocks in turn), then one MMA warp and one copy warp (one lane each). A tile's weights are a fixed stream of
// 17 chunks (wpack_offset): layer 2 = 2 N parts x 4 K chunks, layer 3 = 5 K chunks, layer 4 = 4 K chunks, described by a table in shared memory.
// The copy lane runs free of the layer structure: it refills a stage as soon as the MMAs that read it have completed (three stages; the next layer's
// and the next tile's first chunks included -- weights do not depend on activations). The MMA lane commits every N part to its own mbarrier, so
// layer 2's first part is converted (TMEM -> +bias, ReLU -> bf16 A operand of layer 3) under the MMAs of its second part. (One lane doing both, with
// the chunk geometry recomputed per chunk, spent ~1000 cycles of dependent scalar instructions per chunk -- three times the chunk's MMA time.)
constexpr int DQ_NKC2 = (DQ_K2 + dq_kc(DQ_K2) - 1) / dq_kc(DQ_K2), DQ_NKC3 = (DQ_K3 + dq_kc(DQ_K3) - 1) / dq_kc(DQ_K3), DQ_NKC4 = (DQ_K4 + dq_kc(DQ_K4) - 1) / dq_kc(DQ_K4);       // K chunks per layer (2, 4, 2)
constexpr int DQ_EPI_GROUPS = RLPT_DQN_EPI_GROUPS, DQ_EPI_THREADS = 128 * DQ_EPI_GROUPS, DQ_CHUNKS_PER_TILE = 2 * DQ_NKC2 + DQ_NKC3 + DQ_NKC4;
constexpr int DQ_COMPUTE_THREADS = DQ_EPI_THREADS + 32, DQ_THREADS = DQ_EPI_THREADS + 64;        // epilogue warps + MMA warp (named barrier 1) + copy warp
#ifdef RLPT_DQN_TRACE           // debug build: phase timestamps of CTA 0's second tile (epilogue thread 0 and the issuer lane), printed at kernel end
#define DQ_TR_DECL long long tr_[24]; int ntr_ = 0; const bool tr_on_ = blockIdx.x == 0;
#define DQ_TR(it) do { if (tr_on_ && (it) == 1 && ntr_ < 24) tr_[ntr_++] = clock64(); } while (0)
#define DQ_TR_PRINT(who) do { if (tr_on_ && ntr_ > 1) { printf("%s:", who); for (int i_ = 1; i_ < ntr_; ++i_) printf(" %lld", tr_[i_] - tr_[0]); printf("\n"); } } while (0)
#else
#define DQ_TR_DECL
#define DQ_TR(it) do {} while (0)
#define DQ_TR_PRINT(who) do {} while (0)
#endif

struct ChunkInfo { const __nv_bfloat16* src; int rows, kw, kc, k0, last, part; uint32_t a_off, a_kpad, tmem_col; };
__device__ __forceinline__ ChunkInfo chunk_info(const __nv_bfloat16* w2p, const __nv_bfloat16* w3p, const __nv_bfloat16* w4p, int s) {          // s = position in the tile's stream, 0..16
    ChunkInfo ci; const size_t rep = blockIdx.x % DQ_REPLICAS;       // (strides are in bytes, pointers in bf16)
    if (s < 2 * DQ_NKC2) {
        const int part = s >= DQ_NKC2 ? 1 : 0, kc = s - part * DQ_NKC2, n0 = part ? DQ_L2_SPLIT : 0;
        ci.rows = part ? DQ_N2 - DQ_L2_SPLIT : DQ_L2_SPLIT; ci.kc = kc; ci.k0 = kc * dq_kc(DQ_K2); ci.kw = min(dq_kc(DQ_K2), DQ_K2 - ci.k0); ci.last = kc == DQ_NKC2 - 1; ci.part = part;
        ci.src = w2p + rep * (DQ_W2P_STRIDE / 2) + (size_t)n0 * DQ_K2 + (size_t)ci.rows * ci.k0; ci.a_off = SM_A1; ci.a_kpad = DQ_K2; ci.tmem_col = (uint32_t)n0;
    } else if (s < 2 * DQ_NKC2 + DQ_NKC3) {
        const int kc = s - 2 * DQ_NKC2;
        ci.rows = DQ_N3; ci.kc = kc; ci.k0 = kc * dq_kc(DQ_K3); ci.kw = min(dq_kc(DQ_K3), DQ_K3 - ci.k0); ci.last = kc == DQ_NKC3 - 1; ci.part = 0;
        ci.src = w3p + rep * (DQ_W3P_STRIDE / 2) + (size_t)DQ_N3 * ci.k0; ci.a_off = SM_A2; ci.a_kpad = DQ_K3; ci.tmem_col = DQ_N2;
    } else {
        const int kc = s - 2 * DQ_NKC2 - DQ_NKC3;
        ci.rows = DQ_N4; ci.kc = kc; ci.k0 = kc * dq_kc(DQ_K4); ci.kw = min(dq_kc(DQ_K4), DQ_K4 - ci.k0); ci.last = kc == DQ_NKC4 - 1; ci.part = 0;
        ci.src = w4p + rep * (DQ_W4P_STRIDE / 2) + (size_t)DQ_N4 * ci.k0; ci.a_off = SM_A1; ci.a_kpad = DQ_K4; ci.tmem_col = 0;
    }
    return ci;
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by one thread for the whole CTA; the two shared-memory descriptors are given as
// (low, high) words: the K step only moves the start address (low word, 16-byte units)
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// chunk table entry (shared memory, built once per CTA from chunk_info)
struct ChunkRow { uint32_t src_lo, src_hi, bytes, a_lo, a_hi, b_hi, idesc, tmem_col, n_mma, flags; };        // flags: bit 0 = first chunk of its accumulator, bit 1 = last, bit 2 = N part
constexpr uint32_t SM_TABLE_BYTES = DQ_CHUNKS_PER_TILE * sizeof(ChunkRow);

// 32 (or 16) accumulator columns of one ray: + bias, ReLU, bf16; written as the next layer's A operand and optionally kept feature-major in HBM
template <int NC, bool KEEP>
__device__ __forceinline__ void hidden_block(uint8_t* smem, const uint32_t* r, int c0, const float* bias, uint32_t a_next_off, int k_pad_next, int row,
                                             __nv_bfloat16* keep, int keep_stride, int ray, bool ray_valid, uint32_t* maskw) {
    uint32_t mword = 0;                                                    // relu'(h) of these units as bits (kept for k_dqn_backward: one word instead of 32 activations)
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        __nv_bfloat162 pk[4];
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c0 + 8 * q), b1 = *reinterpret_cast<const float4*>(bias + c0 + 8 * q + 4);       // (16-byte aligned: c0 is a multiple of 16)
        const float bb[8] = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = fmaxf(__uint_as_float(r[8 * q + 2 * j]) + bb[2 * j], 0.f), b = fmaxf(__uint_as_float(r[8 * q + 2 * j + 1]) + bb[2 * j + 1], 0.f);
            pk[j] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<uint4*>(smem + a_next_off + operand_offset(row, c0 + 8 * q, k_pad_next)) = *reinterpret_cast<uint4*>(&pk[0]);
        if (KEEP && keep && ray_valid) {
            __nv_bfloat16* kp = keep + (size_t)(c0 + 8 * q) * keep_stride + ray;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                kp[0] = pk[j].x; kp[keep_stride] = pk[j].y; kp += 2 * (size_t)keep_stride;
                mword |= ((__bfloat16_as_ushort(pk[j].x) & 0x7fffu) ? 1u : 0u) << (8 * q + 2 * j) | ((__bfloat16_as_ushort(pk[j].y) & 0x7fffu) ? 1u : 0u) << (8 * q + 2 * j + 1);
            }
        }
    }
    if (KEEP && maskw && ray_valid) maskw[(size_t)(c0 >> 5) * keep_stride + ray] = mword;
}
// Epilogue of a hidden layer for one epilogue thread: its warpgroup's 32-column blocks (alternating with the other warpgroup's), each as soon
// as the N part that holds it is complete (n_split: first column of the second part, 0 = one part). Bit c of done_par = parity of the
// completions of part_done[c] before this layer.
template <bool KEEP>
__device__ __forceinline__ void hidden_epilogue(uint8_t* smem, uint64_t* part_done, uint32_t done_par, int n_split, uint32_t tmem_lane_addr, uint32_t tmem_col, int n_pad,
                                                const float* bias, uint32_t a_next_off, int k_pad_next, int row, int half, __nv_bfloat16* keep, int keep_stride, int ray, bool ray_valid, uint32_t* maskw) {
    const int n_blocks = (n_pad + 31) / 32;
    for (int b = half; b < n_blocks; b += DQ_EPI_GROUPS) {
        const int c0 = 32 * b, c = (n_split > 0 && c0 >= n_split) ? 1 : 0;
        mbar_wait(&part_done[c], (done_par >> c) & 1u); tc_fence_after();
        uint32_t r[32];
        if (c0 + 32 <= n_pad) {
            tmem_ld32_issue(tmem_lane_addr + tmem_col + (uint32_t)c0, r); tmem_ld_wait();
            hidden_block<32, KEEP>(smem, r, c0, bias, a_next_off, k_pad_next, row, keep, keep_stride, ray, ray_valid, maskw);
        } else {
            tmem_ld16_issue(tmem_lane_addr + tmem_col + (uint32_t)c0, r); tmem_ld_wait();
            hidden_block<16, KEEP>(smem, r, c0, bias, a_next_off, k_pad_next, row, keep, keep_stride, ray, ray_valid, maskw);
        }
    }
}

// KEEP: the training launch (activations and relu' bits of the first batch kept for the backward pass); the inference launches carry none of that code
template <bool KEEP>
__global__ void __launch_bounds__(DQ_THREADS, 1) k_dqn_forward(const __grid_constant__ DqnFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const float4* s_l1c = reinterpret_cast<const float4*>(smem + SM_C1); float4* s_l1 = reinterpret_cast<float4*>(smem + SM_C1);
    float* s_b2 = reinterpret_cast<float*>(smem + SM_BIAS2); float* s_b3 = reinterpret_cast<float*>(smem + SM_BIAS3); float* s_b4 = reinterpret_cast<float*>(smem + SM_BIAS4);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t *b_full = bars, *b_free = bars + DQ_STAGES, *part_done = bars + 2 * DQ_STAGES;       // [3], [3], [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DQ_STAGES + 2);
    const int t = threadIdx.x, warp = t >> 5;

    // layer-1 constants as one float4 (c1, M1 row) per output: the shared-memory pipe takes one instruction per cycle, broadcast or not
    RLPT_PDL_SYNC();
    // Unit DQ_H1 of layer 1 and unit DQ_H2 of layer 2 (padding: their weights are zero in both directions) are constant one: the kept activations then carry
    // the "ones" row that makes the weight-gradient GEMMs of the backward pass produce the bias gradients as well.
    for (int i = t; i < DQ_K2; i += DQ_THREADS) s_l1[i] = i < DQ_H1 ? make_float4(p.c1[i], p.m1[3 * i], p.m1[3 * i + 1], p.m1[3 * i + 2]) : make_float4(i == DQ_H1 ? 1.f : 0.f, 0.f, 0.f, 0.f);
    for (int i = t; i < DQ_N2; i += DQ_THREADS) s_b2[i] = i < DQ_H2 ? p.b2[i] : (i == DQ_H2 ? 1.f : 0.f);
    for (int i = t; i < DQ_N3; i += DQ_THREADS) s_b3[i] = i < DQ_H3 ? p.b3[i] : 0.f;
    for (int i = t; i < DQ_N4; i += DQ_THREADS) s_b4[i] = i < DQ_OUT ? p.b4[i] : 0.f;
    ChunkRow* table = reinterpret_cast<ChunkRow*>(smem + SM_TABLE);
    if (t < DQ_CHUNKS_PER_TILE) {
        const ChunkInfo ci = chunk_info(p.w2p, p.w3p, p.w4p, t);
        const uint32_t a_addr = smem_u32(smem + ci.a_off) + (uint32_t)ci.k0 * 16u;
        ChunkRow cr;
        cr.src_lo = (uint32_t)reinterpret_cast<uint64_t>(ci.src); cr.src_hi = (uint32_t)(reinterpret_cast<uint64_t>(ci.src) >> 32); cr.bytes = (uint32_t)ci.rows * (uint32_t)ci.kw * 2u;
        cr.a_lo = ((a_addr & 0x3FFFFu) >> 4) | ((128u >> 4) << 16); cr.a_hi = ((ci.a_kpad * 16u) >> 4) | (1u << 14); cr.b_hi = (((uint32_t)ci.kw * 16u) >> 4) | (1u << 14);
        cr.idesc = idesc_bf16(DQ_TILE, ci.rows); cr.tmem_col = ci.tmem_col; cr.n_mma = (uint32_t)ci.kw / 16u;
        cr.flags = (ci.kc == 0 ? 1u : 0u) | (ci.last ? 2u : 0u) | ((uint32_t)ci.part << 2);
        table[t] = cr;
    }
    if (t == 0) { for (int i = 0; i < 2 * DQ_STAGES + 2; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t b_lo0 = ((smem_u32(smem + SM_B0) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);       // B descriptor, low word, stage 0
    auto compute_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(DQ_COMPUTE_THREADS) : "memory"); };       // epilogue warps + MMA warp

    const int n_rays1 = p.n_ptr ? min(*p.n_ptr, p.n) : p.n;
    const int n_tiles1 = (n_rays1 + DQ_TILE - 1) / DQ_TILE, n_tiles = n_tiles1 + (p.pos2 ? (p.n2 + DQ_TILE - 1) / DQ_TILE : 0);
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == DQ_EPI_THREADS / 32 + 1) {
        // ---- copy warp (one lane): the weight stream, DQ_STAGES - 1 chunks ahead of the MMAs at most; it takes no part in the layer barriers
        if ((t & 31) == 0) {
            const uint32_t total = (uint32_t)my_tiles * DQ_CHUNKS_PER_TILE;
            uint32_t stage = 0, par = 1, srow = 0;                          // par: parity to wait for on b_free (first pass: the stages have never been used)
            for (uint32_t g = 0; g < total; ++g) {
                if (g >= DQ_STAGES) mbar_wait(&b_free[stage], par);
                const ChunkRow& cr = table[srow];
                const uint32_t bytes = cr.bytes;
                mbar_expect_tx(&b_full[stage], bytes);
                {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(((uint64_t)cr.src_hi << 32) | cr.src_lo);
                    const uint32_t piece = RLPT_DQN_PIECES > 1 ? (((bytes / RLPT_DQN_PIECES) + 127u) & ~127u) : bytes;       // (several bulk copies per chunk: an A/B switch)
                    for (uint32_t o = 0; o < bytes; o += piece) bulk_copy_g2s(smem + SM_B0 + stage * DQ_STAGE_BYTES + o, src + o, min(piece, bytes - o), &b_full[stage]);
                }
                if (++srow == DQ_CHUNKS_PER_TILE) srow = 0;
                if (++stage == DQ_STAGES) { stage = 0; par ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == DQ_EPI_THREADS / 32) {
        // ---- MMA warp: lane 0 works, the warp reconverges before every barrier of the compute group
        const bool lead = (t & 31) == 0;
        DQ_TR_DECL
        uint32_t stage = 0, par = 0, srow = 0;
        auto run = [&](int n) {
            for (int i = 0; i < n; ++i) {
                mbar_wait(&b_full[stage], par);
                tc_fence_after();
                const ChunkRow cr = table[srow];
                uint32_t a_lo = cr.a_lo, b_lo = b_lo0 + stage * (DQ_STAGE_BYTES >> 4);
                const uint32_t d_addr = tmem_base + cr.tmem_col, first = cr.flags & 1u;
                for (uint32_t k = 0; k < cr.n_mma; ++k, a_lo += 16u, b_lo += 16u)
                    umma_bf16_lohi(d_addr, a_lo, cr.a_hi, b_lo, cr.b_hi, cr.idesc, (first ^ 1u) | k);
                umma_commit(&b_free[stage]);
                if (cr.flags & 2u) umma_commit(&part_done[(cr.flags >> 2) & 1u]);
                if (++srow == DQ_CHUNKS_PER_TILE) srow = 0;
                if (++stage == DQ_STAGES) { stage = 0; par ^= 1u; }
            }
        };
        for (int it = 0; it < my_tiles; ++it) {
            if (lead) DQ_TR(it);
            compute_sync();                                                // layer-1 activations are in A1, the previous tile's accumulators have been read
            if (lead) { DQ_TR(it); tc_fence_after(); run(2 * DQ_NKC2); DQ_TR(it); }
            __syncwarp(); compute_sync();                                  // A2 written
            if (lead) { DQ_TR(it); tc_fence_after(); run(DQ_NKC3); DQ_TR(it); }
            __syncwarp(); compute_sync();                                  // A1 (layer-4 operand) written
            if (lead) { DQ_TR(it); tc_fence_after(); run(DQ_NKC4); DQ_TR(it); }
            __syncwarp();
        }
        if (lead) DQ_TR_PRINT("issuer  [sync1 | L2 issued | sync2 | L3 issued | sync3 | L4 issued]");
        __syncwarp();
    } else {
        // ---- epilogue warps
        const int row = t & (DQ_TILE - 1), half = t >> 7;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);        // this warp's quarter of the 128 TMEM lanes
        uint32_t done_par = 0;                                                  // bit c: parity of part_done[c]'s completed phases
        float4 x_next = make_float4(0.f, 0.f, 0.f, 0.f);
        // CTAs that run many tiles start staggered: in lockstep all 148 stream the same weight chunk from L2 and write their Q tiles to HBM at the
        // same moment, and each of those bursts ran at the L2 / HBM limit while the average traffic is a tenth of it
        if (RLPT_DQN_STAGGER && my_tiles >= 4) { const long long t0 = clock64(), d = (long long)(blockIdx.x % 8u) * RLPT_DQN_STAGGER; while (clock64() - t0 < d) { } }
        DQ_TR_DECL
        for (int it = 0; it < my_tiles; ++it) {
            const int tile_all = (int)blockIdx.x + it * (int)gridDim.x;
            // the second batch (next states of a training step: Q only) rides in the same launch, its tiles after the first batch's
            const bool second = tile_all >= n_tiles1;
            const int tile = second ? tile_all - n_tiles1 : tile_all, n_rays = second ? p.n2 : n_rays1;
            const float4* __restrict__ pos = second ? p.pos2 : p.pos;
            float* __restrict__ q_out = second ? p.q2 : p.q; const int q_stride = second ? p.q_stride2 : p.q_stride;
            __nv_bfloat16* const k1 = second ? nullptr : p.h1t; __nv_bfloat16* const k2 = second ? nullptr : p.h2t; __nv_bfloat16* const k3 = second ? nullptr : p.h3t;
            const int ray = tile * DQ_TILE + row; const bool valid = ray < n_rays;
            if (t == 0) DQ_TR(it);
            // ---- layer 1 (fp32): h1 = relu(c1 - M1 x), 8 outputs per 16-byte store into the A operand; the warpgroups take the 26 groups of 8 in turn
            const float4 x = it == 0 ? (valid ? pos[ray] : make_float4(0.f, 0.f, 0.f, 0.f)) : x_next;
            {   // the next tile's position is fetched now: a DRAM round trip at the head of every tile was a tenth of the tile time
                const int ta = tile_all + (int)gridDim.x;
                if (ta < n_tiles) {
                    const bool sec = ta >= n_tiles1; const int r2 = (sec ? ta - n_tiles1 : ta) * DQ_TILE + row;
                    x_next = r2 < (sec ? p.n2 : n_rays1) ? (sec ? p.pos2 : p.pos)[r2] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            for (int j0 = 8 * half; j0 < DQ_K2; j0 += 8 * DQ_EPI_GROUPS) {
                __nv_bfloat162 pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int a = j0 + 2 * j, b = a + 1;
                    const float4 ca = s_l1c[a], cb = s_l1c[b];
                    float ha = fmaxf(ca.x - (ca.y * x.x + ca.z * x.y + ca.w * x.z), 0.f);
                    float hb = fmaxf(cb.x - (cb.y * x.x + cb.z * x.y + cb.w * x.z), 0.f);
                    pk[j] = __floats2bfloat162_rn(ha, hb);
                }
                *reinterpret_cast<uint4*>(smem + SM_A1 + operand_offset(row, j0, DQ_K2)) = *reinterpret_cast<uint4*>(&pk[0]);
                if (KEEP && k1 && valid) {
                    __nv_bfloat16* kp = k1 + (size_t)j0 * p.h_stride + ray;
                    uint32_t mb = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        kp[0] = pk[j].x; kp[p.h_stride] = pk[j].y; kp += 2 * (size_t)p.h_stride;
                        mb |= ((__bfloat16_as_ushort(pk[j].x) & 0x7fffu) ? 1u : 0u) << (2 * j) | ((__bfloat16_as_ushort(pk[j].y) & 0x7fffu) ? 1u : 0u) << (2 * j + 1);
                    }
                    // units j0 .. j0 + 7 = byte (j0 & 31) / 8 of mask word j0 / 32 (with four warpgroups: byte `half` of word number `iteration`)
                    if (p.mask1) reinterpret_cast<uint8_t*>(p.mask1)[((size_t)(j0 >> 5) * p.h_stride + ray) * 4 + ((j0 & 31) >> 3)] = (uint8_t)mb;
                }
            }
            if (t == 0) DQ_TR(it);
            fence_proxy_async(); tc_fence_before(); compute_sync();
            if (t == 0) DQ_TR(it);
            // ---- layer 2: 200 -> 300 (TMEM columns [0, 304), two N parts)
#ifdef RLPT_DQN_TRACE
            if (t == 0) { mbar_wait(&part_done[0], done_par & 1u); DQ_TR(it); }
#endif
            hidden_epilogue<KEEP>(smem, part_done, done_par, DQ_L2_SPLIT, tmem_lane, 0, DQ_N2, s_b2, SM_A2, DQ_K3, row, half, k2, p.h_stride, ray, valid, second ? nullptr : p.mask2);
            done_par ^= 3u;
            if (t == 0) DQ_TR(it);
            fence_proxy_async(); tc_fence_before(); compute_sync();
            if (t == 0) DQ_TR(it);
            // ---- layer 3: 300 -> 200 (columns [304, 512))
#ifdef RLPT_DQN_TRACE
            if (t == 0) { mbar_wait(&part_done[0], done_par & 1u); DQ_TR(it); }
#endif
            hidden_epilogue<KEEP>(smem, part_done, done_par, 0, tmem_lane, DQ_N2, DQ_N3, s_b3, SM_A1, DQ_K4, row, half, k3, p.h_stride, ray, valid, second ? nullptr : p.mask3);
            done_par ^= 1u;
            if (t == 0) DQ_TR(it);
            fence_proxy_async(); tc_fence_before(); compute_sync();
            if (t == 0) DQ_TR(it);
            // ---- layer 4: 200 -> 144 (columns [0, 144)), ReLU on the output as well (N/dq_network.cu:17)
            mbar_wait(&part_done[0], done_par & 1u); tc_fence_after();     // (also: layer 4's MMAs have all read A1 before the next tile's layer 1 rewrites it)
            done_par ^= 1u;
            if (t == 0) DQ_TR(it);
            for (int b = half; b < (DQ_N4 + 31) / 32; b += DQ_EPI_GROUPS) {
                const int c0 = 32 * b;
                uint32_t r[32];
                if (c0 + 32 <= DQ_N4) { tmem_ld32_issue(tmem_lane + (uint32_t)c0, r); tmem_ld_wait(); } else { tmem_ld16_issue(tmem_lane + (uint32_t)c0, r); tmem_ld_wait(); }
                const int nc = min(32, DQ_N4 - c0);
                if (valid) {
                    float* qp = q_out + (size_t)c0 * q_stride + ray;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) if (4 * j4 < nc) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_b4 + c0 + 4 * j4);
                        qp[0] = fmaxf(__uint_as_float(r[4 * j4]) + b4.x, 0.f); qp[q_stride] = fmaxf(__uint_as_float(r[4 * j4 + 1]) + b4.y, 0.f);
                        qp[2 * (size_t)q_stride] = fmaxf(__uint_as_float(r[4 * j4 + 2]) + b4.z, 0.f); qp[3 * (size_t)q_stride] = fmaxf(__uint_as_float(r[4 * j4 + 3]) + b4.w, 0.f);
                        qp += 4 * (size_t)q_stride;
                    }
                }
            }
            tc_fence_before();                                              // (the next tile's first CTA barrier orders these TMEM reads before its MMAs)
            if (t == 0) DQ_TR(it);
        }
        if (t == 0) DQ_TR_PRINT("epilogue [L1 done | sync1 | L2 part0 ready | E2 done | sync2 | L3 ready | E3 done | sync3 | L4 ready | E4 done]");
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int dqn_gemm_set_smem_limit();
__global__ void k_dqn_backward(const __grid_constant__ DqnBwdParams p);
int dqn_set_smem_limit() {
    int rc = (int)cudaFuncSetAttribute(k_dqn_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
    if (!rc) rc = (int)cudaFuncSetAttribute(k_dqn_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
    if (!rc) rc = (int)cudaFuncSetAttribute(k_dqn_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
    return rc ? rc : dqn_gemm_set_smem_limit();
}

int dqn_forward(const DqnDev& d, const DqnFwdParams& p, cudaStream_t s) {
    if (!d.ready || p.n <= 0) return p.n == 0 ? 0 : -1;        // p.n bounds the launch; p.n_ptr (if set) gives the live count
    int dev = 0, n_sm = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = (p.n + DQ_TILE - 1) / DQ_TILE + (p.pos2 ? (p.n2 + DQ_TILE - 1) / DQ_TILE : 0);
    if (p.h1t) k_dqn_forward<true><<<n_tiles < n_sm ? n_tiles : n_sm, DQ_THREADS, SM_TOTAL, s>>>(p);
    else k_dqn_forward<false><<<n_tiles < n_sm ? n_tiles : n_sm, DQ_THREADS, SM_TOTAL, s>>>(p);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ backward data path, one kernel
// The TD step's backward data path for a tile of 128 rays, in k_dqn_forward's own shape (same warp roles, chunk stream, stages, barriers):
//   phase 0   TD target (max over the next state's Q cos(theta), four threads per ray), output-layer gradient g, delta3 = g W4[a, :] relu'(h3) written as
//             the bf16 A operand (and feature-major to HBM for the weight-gradient GEMM), dW4 / db4 / loss scatter-added
//   P2        delta3 [128 x 208] x (W3^T)^T on the tensor cores (W3^T packed like the forward's layer 2: two N parts x four K chunks) -> TMEM [0, 304)
//             epilogue: relu'(h2) mask -> bf16 A operand of the next product + d2t in HBM
//   P1        delta2 [128 x 304] x (W2^T)^T (packed like the forward's layer 3) -> TMEM [304, 512); epilogue: relu'(h1) mask -> d1t in HBM
// It replaces k_delta3, two data GEMMs and two mask kernels (43 us of a 92 us optimiser step) -- the activations never leave the SM between the products.
constexpr int DQ_BWD_CHUNKS = 2 * DQ_NKC2 + DQ_NKC3;
static_assert(DQ_REPLICAS == 1, "the packed transposes of k_dqn_backward have no replicas");
static_assert(DQ_EPI_GROUPS == 4, "mask bytes of layer 1 / phase 0 of k_dqn_backward: four warpgroups, eight units each per 32-unit word");
template <int NC>
__device__ __forceinline__ void bwd_block(uint8_t* smem, const uint32_t* r, uint32_t mbits, int c0, bool write_a, uint32_t a_next_off, int k_pad_next, int row, __nv_bfloat16* out, int S, int ray) {
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        __nv_bfloat162 pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = ((mbits >> (8 * q + 2 * j)) & 1u) ? __uint_as_float(r[8 * q + 2 * j]) : 0.f, b = ((mbits >> (8 * q + 2 * j + 1)) & 1u) ? __uint_as_float(r[8 * q + 2 * j + 1]) : 0.f;
            pk[j] = __floats2bfloat162_rn(a, b);
        }
        if (write_a) *reinterpret_cast<uint4*>(smem + a_next_off + operand_offset(row, c0 + 8 * q, k_pad_next)) = *reinterpret_cast<uint4*>(&pk[0]);
        __nv_bfloat16* op = out + (size_t)(c0 + 8 * q) * S + ray;
#pragma unroll
        for (int j = 0; j < 4; ++j) { op[0] = pk[j].x; op[S] = pk[j].y; op += 2 * (size_t)S; }
    }
}
__device__ __forceinline__ void bwd_epilogue(uint8_t* smem, uint64_t* part_done, uint32_t done_par, int n_split, uint32_t tmem_lane_addr, uint32_t tmem_col, int n_pad, int n_feat,
                                             const uint32_t* __restrict__ maskw, bool write_a, uint32_t a_next_off, int k_pad_next, int row, int grp, __nv_bfloat16* out, int S, int ray) {
    const int n_blocks = (n_pad + 31) / 32;
    for (int b = grp; b < n_blocks; b += DQ_EPI_GROUPS) {
        const int c0 = 32 * b, c = (n_split > 0 && c0 >= n_split) ? 1 : 0, nc = min(32, n_pad - c0);
        // relu'(h) of this block's units: one word the forward pass left behind (requested before the product is waited for); padding units masked off
        const int nf = n_feat - c0;
        const uint32_t mbits = __ldg(maskw + (size_t)b * S + ray) & (nf >= 32 ? 0xffffffffu : (nf > 0 ? (1u << nf) - 1u : 0u));
        mbar_wait(&part_done[c], (done_par >> c) & 1u); tc_fence_after();
        uint32_t r[32];
        if (nc == 32) { tmem_ld32_issue(tmem_lane_addr + tmem_col + (uint32_t)c0, r); tmem_ld_wait(); bwd_block<32>(smem, r, mbits, c0, write_a, a_next_off, k_pad_next, row, out, S, ray); }
        else { tmem_ld16_issue(tmem_lane_addr + tmem_col + (uint32_t)c0, r); tmem_ld_wait(); bwd_block<16>(smem, r, mbits, c0, write_a, a_next_off, k_pad_next, row, out, S, ray); }
    }
}
__global__ void __launch_bounds__(DQ_THREADS, 1) k_dqn_backward(const __grid_constant__ DqnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_best = reinterpret_cast<float*>(smem + SM_C1);               // [4][128] partial maxima of the TD target
    float* s_g = reinterpret_cast<float*>(smem + SM_BIAS2); int* s_a = reinterpret_cast<int*>(smem + SM_BIAS2 + 4 * DQ_TILE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t *b_full = bars, *b_free = bars + DQ_STAGES, *part_done = bars + 2 * DQ_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DQ_STAGES + 2);
    const int t = threadIdx.x, warp = t >> 5;
    RLPT_PDL_SYNC();
    ChunkRow* table = reinterpret_cast<ChunkRow*>(smem + SM_TABLE);
    if (t < DQ_BWD_CHUNKS) {
        const ChunkInfo ci = chunk_info(p.w3tp, p.w2tp, nullptr, t);
        const uint32_t a_addr = smem_u32(smem + ci.a_off) + (uint32_t)ci.k0 * 16u;
        ChunkRow cr;
        cr.src_lo = (uint32_t)reinterpret_cast<uint64_t>(ci.src); cr.src_hi = (uint32_t)(reinterpret_cast<uint64_t>(ci.src) >> 32); cr.bytes = (uint32_t)ci.rows * (uint32_t)ci.kw * 2u;
        cr.a_lo = ((a_addr & 0x3FFFFu) >> 4) | ((128u >> 4) << 16); cr.a_hi = ((ci.a_kpad * 16u) >> 4) | (1u << 14); cr.b_hi = (((uint32_t)ci.kw * 16u) >> 4) | (1u << 14);
        cr.idesc = idesc_bf16(DQ_TILE, ci.rows); cr.tmem_col = ci.tmem_col; cr.n_mma = (uint32_t)ci.kw / 16u;
        cr.flags = (ci.kc == 0 ? 1u : 0u) | (ci.last ? 2u : 0u) | ((uint32_t)ci.part << 2);
        table[t] = cr;
    }
    if (t == 0) { for (int i = 0; i < 2 * DQ_STAGES + 2; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t b_lo0 = ((smem_u32(smem + SM_B0) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
    auto compute_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(DQ_COMPUTE_THREADS) : "memory"); };       // epilogue warps + MMA warp
    auto epi_sync = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(DQ_EPI_THREADS) : "memory"); };               // epilogue warps only
    const int S = p.S, n_tiles = S / DQ_TILE;
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == DQ_EPI_THREADS / 32 + 1) {
        if ((t & 31) == 0) {
            const uint32_t total = (uint32_t)my_tiles * DQ_BWD_CHUNKS;
            uint32_t stage = 0, par = 1, srow = 0;
            for (uint32_t g = 0; g < total; ++g) {
                if (g >= DQ_STAGES) mbar_wait(&b_free[stage], par);
                const ChunkRow& cr = table[srow];
                const uint32_t bytes = cr.bytes;
                mbar_expect_tx(&b_full[stage], bytes);
                {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(((uint64_t)cr.src_hi << 32) | cr.src_lo);
                    const uint32_t piece = RLPT_DQN_PIECES > 1 ? (((bytes / RLPT_DQN_PIECES) + 127u) & ~127u) : bytes;       // (several bulk copies per chunk: an A/B switch)
                    for (uint32_t o = 0; o < bytes; o += piece) bulk_copy_g2s(smem + SM_B0 + stage * DQ_STAGE_BYTES + o, src + o, min(piece, bytes - o), &b_full[stage]);
                }
                if (++srow == DQ_BWD_CHUNKS) srow = 0;
                if (++stage == DQ_STAGES) { stage = 0; par ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == DQ_EPI_THREADS / 32) {
        const bool lead = (t & 31) == 0;
        uint32_t stage = 0, par = 0, srow = 0;
        auto run = [&](int n) {
            for (int i = 0; i < n; ++i) {
                mbar_wait(&b_full[stage], par);
                tc_fence_after();
                const ChunkRow cr = table[srow];
                uint32_t a_lo = cr.a_lo, b_lo = b_lo0 + stage * (DQ_STAGE_BYTES >> 4);
                const uint32_t d_addr = tmem_base + cr.tmem_col, first = cr.flags & 1u;
                for (uint32_t k = 0; k < cr.n_mma; ++k, a_lo += 16u, b_lo += 16u)
                    umma_bf16_lohi(d_addr, a_lo, cr.a_hi, b_lo, cr.b_hi, cr.idesc, (first ^ 1u) | k);
                umma_commit(&b_free[stage]);
                if (cr.flags & 2u) umma_commit(&part_done[(cr.flags >> 2) & 1u]);
                if (++srow == DQ_BWD_CHUNKS) srow = 0;
                if (++stage == DQ_STAGES) { stage = 0; par ^= 1u; }
            }
        };
        for (int it = 0; it < my_tiles; ++it) {
            compute_sync();                                                // delta3 is in A1
            if (lead) { tc_fence_after(); run(2 * DQ_NKC2); }
            __syncwarp(); compute_sync();                                  // delta2 is in A2
            if (lead) { tc_fence_after(); run(DQ_NKC3); }
            __syncwarp();
        }
    } else {
        const int row = t & (DQ_TILE - 1), grp = t >> 7;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t done_par = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x, ray = tile * DQ_TILE + row; const bool valid = ray < p.n;
            // ---- phase 0: TD target, output-layer gradient, delta3
            if (p.td.q_next) {
                float best = 0.f;
                if (valid) {
#pragma unroll 12
                    for (int k = grp * (DQ_OUT / DQ_EPI_GROUPS); k < (grp + 1) * (DQ_OUT / DQ_EPI_GROUPS); ++k) best = fmaxf(best, __ldg(p.td.q_next + (size_t)k * p.td.q_stride + ray) * c_dq_cos[k]);
                }
                s_best[grp * DQ_TILE + row] = best;
                epi_sync();
            }
            if (grp == 0) {
                float g = 0.f, loss = 0.f; int a = 0;
                if (valid) {
                    float target;
                    if (p.td.q_next) {
                        float best = 0.f;
#pragma unroll
                        for (int w = 0; w < DQ_EPI_GROUPS; ++w) best = fmaxf(best, s_best[w * DQ_TILE + row]);
                        target = p.td.reward[ray];
                        if (p.td.state[ray] != 1u) target += best * p.td.discount[ray];
                        p.targets[ray] = target;
                    } else target = p.targets[ray];
                    a = (int)p.actions[ray]; const float qa = p.q[(size_t)a * S + ray], diff = qa - target;
                    loss = diff * diff; g = qa > 0.f ? 2.f * diff : 0.f;
                    if (g != 0.f) atomicAdd(p.gb4 + a, g);
                }
                s_g[row] = g; s_a[row] = a; p.g_out[ray] = g;
                for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
                if ((t & 31) == 0 && loss != 0.f) atomicAdd(p.scalars, loss);
            }
            epi_sync();
            {
                const float g = s_g[row]; const int a = s_a[row];
                // relu'(h3) as mask words (unit j0 + e = bit 8 grp + e of word number j0 / 32): seven loads per thread, all in flight together
                uint32_t m3[(DQ_K4 + 31) / 32];
#pragma unroll
                for (int w = 0; w < (DQ_K4 + 31) / 32; ++w) m3[w] = valid ? __ldg(p.mask3 + (size_t)w * S + ray) : 0u;
#pragma unroll
                for (int it = 0; it < (DQ_K4 + 31) / 32; ++it) {
                    const int j0 = 8 * grp + 32 * it;
                    if (j0 >= DQ_K4) break;
                    float w[8];
                    const uint32_t hbits = j0 < DQ_H3 ? (m3[it] >> (8 * grp)) & 0xffu : 0u;
                    if (g != 0.f && j0 < DQ_H3) {
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w4 + (size_t)a * DQ_H3 + j0)), w1 = __ldg(reinterpret_cast<const float4*>(p.w4 + (size_t)a * DQ_H3 + j0 + 4));
                        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[e] = 0.f;
                    }
                    __nv_bfloat162 pk[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        const bool on0 = g != 0.f && ((hbits >> e) & 1u), on1 = g != 0.f && ((hbits >> (e + 1)) & 1u);
                        pk[e >> 1] = __floats2bfloat162_rn(on0 ? g * w[e] : 0.f, on1 ? g * w[e + 1] : 0.f);       // (dW4's scatter-adds: k_dw4_rank1 beside the GEMMs -- from these 32 SMs they took 45 us)
                    }
                    *reinterpret_cast<uint4*>(smem + SM_A1 + operand_offset(row, j0, DQ_K4)) = *reinterpret_cast<uint4*>(&pk[0]);
                    __nv_bfloat16* op = p.d3t + (size_t)j0 * S + ray;
#pragma unroll
                    for (int e = 0; e < 4; ++e) { op[0] = pk[e].x; op[S] = pk[e].y; op += 2 * (size_t)S; }
                }
            }
            fence_proxy_async(); tc_fence_before(); compute_sync();
            // ---- P2 = delta3 W3 -> relu'(h2) -> delta2 (A operand + d2t)
            bwd_epilogue(smem, part_done, done_par, DQ_L2_SPLIT, tmem_lane, 0, DQ_N2, DQ_H2, p.mask2, true, SM_A2, DQ_K3, row, grp, p.d2t, S, ray);
            done_par ^= 3u;
            fence_proxy_async(); tc_fence_before(); compute_sync();
            // ---- P1 = delta2 W2 -> relu'(h1) -> d1t
            bwd_epilogue(smem, part_done, done_par, 0, tmem_lane, DQ_N2, DQ_N3, DQ_H1, p.mask1, false, 0, 0, row, grp, p.d1t, S, ray);
            done_par ^= 1u;
            tc_fence_before();
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}
int dqn_backward(const DqnBwdParams& p, cudaStream_t s) {
    int dev = 0, n_sm = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = p.S / DQ_TILE;
    k_dqn_backward<<<n_tiles < n_sm ? n_tiles : n_sm, DQ_THREADS, SM_TOTAL, s>>>(p);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ operands from parameters
// one thread per packed element: bf16 copy of W [n][k] (row-major fp32, n x k) into the canonical [n_pad][k_pad] operand, zero padding
__global__ void k_pack_weights(const float* __restrict__ w, int n, int k, int n_pad, int k_pad, int n_split, __nv_bfloat16* __restrict__ out, size_t rep_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * k_pad) return;
    int row = i / k_pad, col = i % k_pad;
    float v = (row < n && col < k) ? w[(size_t)row * k + col] : 0.f;
    const size_t off = wpack_offset(n_split, row, col, n_pad, k_pad);
    for (int r = 0; r < DQ_REPLICAS; ++r) *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(out) + r * rep_stride + off) = __float2bfloat16_rn(v);
}
// c1 = b1 + W1 v, M1[:, d] = sum_{i % 3 == d} W1[:, i]; one warp per output row, fp32 with pairwise lane sums
__global__ void k_layer1_operands(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ v, int k_in, float* __restrict__ c1, float* __restrict__ m1) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= DQ_H1) return;
    float acc = 0.f, m[3] = { 0.f, 0.f, 0.f };
    for (int i = lane; i < k_in; i += 32) { float w = w1[(size_t)row * k_in + i]; acc += w * v[i]; int d = i % 3; m[0] += d == 0 ? w : 0.f; m[1] += d == 1 ? w : 0.f; m[2] += d == 2 ? w : 0.f; }
    for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); for (int d = 0; d < 3; ++d) m[d] += __shfl_xor_sync(0xffffffffu, m[d], o); }
    if (lane == 0) { c1[row] = b1[row] + acc; m1[3 * row] = m[0]; m1[3 * row + 1] = m[1]; m1[3 * row + 2] = m[2]; }
}

int dqn_refresh_operands(DqnDev& d, cudaStream_t s) {
    k_layer1_operands<<<(DQ_H1 + 7) / 8, 256, 0, s>>>(d.w[0], d.b[0], d.vertices, d.k_in, d.c1, d.m1);
    k_pack_weights<<<(DQ_N2 * DQ_K2 + 255) / 256, 256, 0, s>>>(d.w[1], DQ_H2, DQ_H1, DQ_N2, DQ_K2, DQ_L2_SPLIT, d.w2p, DQ_W2P_STRIDE);
    k_pack_weights<<<(DQ_N3 * DQ_K3 + 255) / 256, 256, 0, s>>>(d.w[2], DQ_H3, DQ_H2, DQ_N3, DQ_K3, 0, d.w3p, DQ_W3P_STRIDE);
    k_pack_weights<<<(DQ_N4 * DQ_K4 + 255) / 256, 256, 0, s>>>(d.w[3], DQ_OUT, DQ_H3, DQ_N4, DQ_K4, 0, d.w4p, DQ_W4P_STRIDE);
    return (int)cudaGetLastError();
}

void dqn_free(DqnDev& d) {
    for (int l = 0; l < 4; ++l) { cudaFree(d.w[l]); cudaFree(d.b[l]); d.w[l] = d.b[l] = nullptr; }
    cudaFree(d.vertices); cudaFree(d.c1); cudaFree(d.m1); cudaFree(d.w2p); cudaFree(d.w3p); cudaFree(d.w4p);
    d = DqnDev{};
}
int dqn_alloc(DqnDev& d, int k_in) {
    dqn_free(d);
    d.k_in = k_in;
    DqnHost shape; shape.k_in = k_in;
    for (int l = 0; l < 4; ++l) { DQ_CK(cudaMalloc(&d.w[l], sizeof(float) * (size_t)DqnHost::rows(l) * shape.cols(l))); DQ_CK(cudaMalloc(&d.b[l], sizeof(float) * DqnHost::rows(l))); }
    DQ_CK(cudaMalloc(&d.vertices, sizeof(float) * (size_t)k_in)); DQ_CK(cudaMalloc(&d.c1, sizeof(float) * DQ_H1)); DQ_CK(cudaMalloc(&d.m1, sizeof(float) * DQ_H1 * 3));
    DQ_CK(cudaMalloc(&d.w2p, DQ_REPLICAS * DQ_W2P_STRIDE)); DQ_CK(cudaMalloc(&d.w3p, DQ_REPLICAS * DQ_W3P_STRIDE)); DQ_CK(cudaMalloc(&d.w4p, DQ_REPLICAS * DQ_W4P_STRIDE));
    return 0;
}
int dqn_upload(DqnDev& d, const DqnHost& h, const float* vertices, cudaStream_t s) {
    if (d.k_in != h.k_in || !d.w[0]) { int rc = dqn_alloc(d, h.k_in); if (rc) return rc; }
    for (int l = 0; l < 4; ++l) {
        DQ_CK(cudaMemcpyAsync(d.w[l], h.w[l].data(), sizeof(float) * h.w[l].size(), cudaMemcpyHostToDevice, s));
        DQ_CK(cudaMemcpyAsync(d.b[l], h.b[l].data(), sizeof(float) * h.b[l].size(), cudaMemcpyHostToDevice, s));
    }
    DQ_CK(cudaMemcpyAsync(d.vertices, vertices, sizeof(float) * (size_t)h.k_in, cudaMemcpyHostToDevice, s));
    DQ_CK(cudaStreamSynchronize(s));
    int rc = dqn_refresh_operands(d, s); if (rc) return rc;
    d.ready = true;
    return 0;
}
int dqn_download(const DqnDev& d, DqnHost& h, cudaStream_t s) {
    h.k_in = d.k_in;
    for (int l = 0; l < 4; ++l) {
        h.w[l].resize((size_t)DqnHost::rows(l) * h.cols(l)); h.b[l].resize(DqnHost::rows(l));
        DQ_CK(cudaMemcpyAsync(h.w[l].data(), d.w[l], sizeof(float) * h.w[l].size(), cudaMemcpyDeviceToHost, s));
        DQ_CK(cudaMemcpyAsync(h.b[l].data(), d.b[l], sizeof(float) * h.b[l].size(), cudaMemcpyDeviceToHost, s));
    }
    DQ_CK(cudaStreamSynchronize(s));
    return 0;
}

// ------------------------------------------------------------------------------------------------ host: init + DyNet text files
// DyNet's default initialiser (ParameterInitGlorot, DyNet 2.x -- the reference pins no version): uniform in
// [-s, s], s = sqrt(3 * n_dims) / sqrt(sum of dims): sqrt(6 / (rows + cols)) for a matrix, sqrt(3 / rows) for a bias vector.
// The random stream is this library's own (splitmix64); training dynamics are "parity unpinned" by construction.
void dqn_init_glorot(DqnHost& h, int k_in, uint32_t seed) {
    h.k_in = k_in;
    uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(seed + 1u);
    auto next = [&]() { st += 0x9E3779B97F4A7C15ull; uint64_t z = st; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31; return (float)((z >> 40) * (1.0 / 16777216.0)); };
    for (int l = 0; l < 4; ++l) {
        const int r = DqnHost::rows(l), c = h.cols(l);
        h.w[l].resize((size_t)r * c); h.b[l].resize(r);
        const float sw = std::sqrt(6.f / (float)(r + c)), sb = std::sqrt(3.f / (float)r);
        for (float& x : h.w[l]) x = (2.f * next() - 1.f) * sw;
        for (float& x : h.b[l]) x = (2.f * next() - 1.f) * sb;
    }
}

// DyNet TextFileSaver: per parameter "#Parameter# /_i {rows,cols} nbytes ZERO_GRAD\n" then rows*cols values "%+.8e" separated by
// spaces, column-major (Radiance_Map_Data/*.model; G/deep_learning/neural_q_pathtracer.cu:55-59,191-196)
int dqn_load_text(DqnHost& h, const char* path, std::string& err) {
    std::ifstream in(path);
    if (!in.is_open()) { err = std::string("cannot open ") + path; return 1; }
    std::string header, values;
    for (int blk = 0; blk < 8; ++blk) {
        if (!std::getline(in, header) || !std::getline(in, values)) { err = "truncated model file (expected 8 parameter blocks)"; return 1; }
        int rows = 0, cols = 1; char name[64];
        if (header.rfind("#Parameter#", 0) != 0) { err = "not a DyNet text model: " + header.substr(0, 40); return 1; }
        const size_t lb = header.find('{'), rb = header.find('}');
        if (lb == std::string::npos || rb == std::string::npos) { err = "bad header: " + header; return 1; }
        (void)name;
        if (sscanf(header.c_str() + lb, "{%d,%d}", &rows, &cols) < 1) { err = "bad dims: " + header; return 1; }
        const int l = blk / 2; const bool is_w = blk % 2 == 0;
        if (rows != DqnHost::rows(l)) { err = "unexpected layer shape in " + header; return 1; }
        if (is_w && l == 0) h.k_in = cols;
        if (is_w && cols != h.cols(l)) { err = "unexpected layer shape in " + header; return 1; }
        std::vector<float>& dst = is_w ? h.w[l] : h.b[l];
        const int c = is_w ? cols : 1;
        dst.assign((size_t)rows * c, 0.f);
        const char* sp = values.c_str(); char* end = nullptr;
        for (int j = 0; j < c; ++j) for (int i = 0; i < rows; ++i) {            // column-major in the file
            float v = strtof(sp, &end);
            if (end == sp) { err = "too few values in block " + std::to_string(blk); return 1; }
            dst[(size_t)i * c + j] = v; sp = end;
        }
    }
    return 0;
}
int dqn_save_text(const DqnHost& h, const char* path, std::string& err) {
    FILE* f = fopen(path, "w");
    if (!f) { err = std::string("cannot open ") + path; return 1; }
    for (int blk = 0; blk < 8; ++blk) {
        const int l = blk / 2; const bool is_w = blk % 2 == 0;
        const int rows = DqnHost::rows(l), c = is_w ? h.cols(l) : 1;
        const std::vector<float>& src = is_w ? h.w[l] : h.b[l];
        const long nbytes = (long)rows * c * 16 + 1;                               // 15 characters + separator per value, + newline (DyNet writes the byte count)
        if (is_w) fprintf(f, "#Parameter# /_%d {%d,%d} %ld ZERO_GRAD\n", blk, rows, c, nbytes); else fprintf(f, "#Parameter# /_%d {%d} %ld ZERO_GRAD\n", blk, rows, nbytes);
        for (int j = 0; j < c; ++j) for (int i = 0; i < rows; ++i) fprintf(f, "%+.8e ", src[(size_t)i * c + j]);
        fprintf(f, "\n");
    }
    fclose(f);
    return 0;
}

}  // namespace rlpt

// ================================================================================================ training
namespace rlpt {

// ------------------------------------------------------------------------------------------------ generic tcgen05 GEMM
// C[M x N] (+)= A[M x K] * B[N x K]^T; A, B bf16 with K contiguous (row strides lda, ldb, multiples of 8), C fp32 row-major.
// CTA tile 128 x BN (BN <= 160, multiple of 16); grid.z splits K into pieces of at most GM_KMAX, partial tiles are combined with atomicAdd
// (C zeroed by the caller) -- used for the backward pass, where M or K is the batch.
// Single shot: a CTA's whole K piece of both operands (<= 128 x 304 + 160 x 304 bf16 = 175 KB) goes into shared memory with one wave of 16-byte
// cp.async copies (zero-filled past the edges), one thread then issues all K / 16 MMAs, and 8 warps read the accumulator out of TMEM. The backward
// pass is a chain of these on <= 64 CTAs each: what counts is latency, and this form pays the global-memory round trip once instead of once per
// K chunk (the first form staged 64-wide chunks through registers: 15-18 us per launch).
constexpr int GM_BN = 160, GM_KMAX = 304, GM_THREADS = 256;
constexpr uint32_t GM_A = 0, GM_B = GM_A + DQ_TILE * GM_KMAX * 2, GM_BAR = GM_B + GM_BN * GM_KMAX * 2, GM_TOTAL = GM_BAR + 64;
static_assert(GM_TOTAL <= 227 * 1024, "shared-memory budget");
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

// c_t != 0: C is stored transposed, C[n][m] with row length ldc -- a warp's 32 threads (one output row each) then write consecutive addresses,
// and the consumer of the backward data path (k_delta_hidden, feature-major) reads them the same way.
__global__ void __launch_bounds__(GM_THREADS, 1) k_gemm_bf16_tn(const __nv_bfloat16* __restrict__ A, int lda, const __nv_bfloat16* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                                 int M, int N, int K, int k_per_split, int c_t, int bn_tile) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + GM_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x, warp = t >> 5;
    const int m0 = blockIdx.x * DQ_TILE, n0 = blockIdx.y * bn_tile, bn = min(bn_tile, N - n0);
    const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
    const int kk = (k_end - k_begin + 15) & ~15, pk = kk >> 3;            // this CTA's K, padded to the MMA's K step; 16-byte pieces per row
    uint8_t* sa = smem + GM_A; uint8_t* sb = smem + GM_B;
    RLPT_PDL_SYNC();
    for (int idx = t; idx < DQ_TILE * pk; idx += GM_THREADS) {
        const int row = idx / pk, j = idx - row * pk, m = m0 + row, k = k_begin + 8 * j;
        const bool ok = m < M && k < k_end;
        cp_async16(sa + operand_offset(row, 8 * j, kk), ok ? A + (size_t)m * lda + k : A, ok);
    }
    for (int idx = t; idx < bn * pk; idx += GM_THREADS) {
        const int row = idx / pk, j = idx - row * pk, k = k_begin + 8 * j;
        const bool ok = k < k_end;
        cp_async16(sb + operand_offset(row, 8 * j, kk), ok ? B + (size_t)(n0 + row) * ldb + k : B, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (t == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (t == 0 && kk > 0) {
        const uint32_t idesc = idesc_bf16(DQ_TILE, bn);
        uint32_t a_lo = ((smem_u32(sa) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16), b_lo = ((smem_u32(sb) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
        const uint32_t hi = (uint32_t)kk | (1u << 14);                     // SBO = kk * 16 bytes, in 16-byte units
        for (int k = 0; k < kk / 16; ++k, a_lo += 16u, b_lo += 16u) umma_bf16_lohi(tmem_base, a_lo, hi, b_lo, hi, idesc, (uint32_t)k);
        umma_commit(bar);
    }
    if (kk > 0) {
        mbar_wait(bar, 0); tc_fence_after();
        const int m = m0 + (t & (DQ_TILE - 1)); const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const bool split = gridDim.z > 1;
        for (int c0 = 16 * (warp >> 2); c0 < bn; c0 += 32) {               // the two warpgroups take the 16-column blocks in turn
            uint32_t r[16]; tmem_ld16(lane_addr + (uint32_t)c0, r);
            if (m < M) {
                float* dst = c_t ? C + (size_t)(n0 + c0) * ldc + m : C + (size_t)m * ldc + n0 + c0;
                const size_t step = c_t ? (size_t)ldc : 1;
                if (split && !c_t) {                                         // 16 consecutive floats of one row: four 16-byte vector reductions instead of sixteen scalar ones
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { if (split) atomicAdd(dst + j * step, __uint_as_float(r[j])); else dst[j * step] = __uint_as_float(r[j]); }
                }
            }
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
static int gemm_tn(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, float* C, int ldc, int M, int N, int K, int k_splits, cudaStream_t s, int c_t = 0, bool pdl = false, int bn_tile = GM_BN) {
    const int k_cap = 256;                                                 // K per CTA when K has to be split (<= GM_KMAX)
    if (k_splits < (K + GM_KMAX - 1) / GM_KMAX) k_splits = (K + k_cap - 1) / k_cap;
    int k_per = ((K + k_splits - 1) / k_splits + 15) / 16 * 16; if (k_per > GM_KMAX) k_per = k_cap; if (k_per < 16) k_per = 16;
    const int zs = (K + k_per - 1) / k_per;
    dim3 grid((M + DQ_TILE - 1) / DQ_TILE, (N + bn_tile - 1) / bn_tile, zs);
    return (int)launch_ex(pdl, k_gemm_bf16_tn, grid, dim3(GM_THREADS), GM_TOTAL, s, A, lda, B, ldb, C, ldc, M, N, K, k_per, c_t, bn_tile);
}

// ------------------------------------------------------------------------------------------------ element-wise pieces
__global__ void k_transpose_bf16(const float* __restrict__ w, int rows, int cols, int out_rows_pad, int out_cols_pad, __nv_bfloat16* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;                       // out[c][r] = w[r][c], zero padded to [out_rows_pad][out_cols_pad]
    if (i >= out_rows_pad * out_cols_pad) return;
    int c = i / out_cols_pad, r = i % out_cols_pad;
    out[i] = __float2bfloat16_rn((c < cols && r < rows) ? w[(size_t)r * cols + c] : 0.f);
}
// (k_step_begin, below the parameter segment table, prepares a step: zeroing + the batch inputs of the gradient GEMMs)
// output-layer delta (one action per ray): g = d/dq (target - q_a)^2 * relu'(q_a); delta3 = g W4[a, :] relu'(h3);
// db4[a] += g; loss accumulated (dW4: k_dw4_rank1). One CTA per 32 rays x all 208 hidden units: the feature-major arrays (h3t, d3t) are
// walked with the ray index fastest, the ray-major d3 is written from a shared-memory tile with the unit index fastest -- every access
// coalesced. (One thread per ray walking 208 units, the first form of this kernel, took 125 us of a 270 us optimiser step.)
constexpr int D3_RAYS = 32;
// With tdp.q_next set the kernel first derives the batch's TD targets itself (compute_td_targets, nn_rendering_helpers.cu:91-140:
// reward + discount * max_a Q(s', a) cos(theta_a); terminal: the reward) -- eight warps x 18 cells per ray, combined through shared memory --
// instead of reading them from a kernel of its own (8 us of a 130 us optimiser step).
__global__ void __launch_bounds__(256) k_delta3(const float* __restrict__ q, const uint32_t* __restrict__ actions, float* __restrict__ targets, DqnTdParams tdp, int n, int S,
                                                const float* __restrict__ w4, const __nv_bfloat16* __restrict__ h3t, __nv_bfloat16* __restrict__ d3, __nv_bfloat16* __restrict__ d3t,
                                                float* __restrict__ g_out, float* __restrict__ gb4, float* __restrict__ scalars) {
    __shared__ __nv_bfloat16 tile[D3_RAYS][DQ_K4 + 2];                   // + 2: rows 105 words apart, conflict-free column writes
    __shared__ float s_g[D3_RAYS]; __shared__ int s_a[D3_RAYS]; __shared__ float s_best[8][D3_RAYS];
    const int i0 = blockIdx.x * D3_RAYS, tid = threadIdx.x;
    const int ii = tid & 31, jj = tid >> 5, i = i0 + ii;
    RLPT_PDL_SYNC();
    if (tdp.q_next) {
        float best = 0.f;
        if (i < n) {
#pragma unroll
            for (int k = jj * 18; k < jj * 18 + 18; ++k) best = fmaxf(best, __ldg(tdp.q_next + (size_t)k * tdp.q_stride + i) * c_dq_cos[k]);
        }
        s_best[jj][ii] = best;
        __syncthreads();
    }
    if (tid < D3_RAYS) {
        float g = 0.f, loss = 0.f; int a = 0;
        if (i < n) {
            float target;
            if (tdp.q_next) {
                float best = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) best = fmaxf(best, s_best[w][ii]);
                target = tdp.reward[i];
                if (tdp.state[i] != 1u) target += best * tdp.discount[i];
                targets[i] = target;
            } else target = targets[i];
            a = (int)actions[i]; const float qa = q[(size_t)a * S + i], diff = qa - target;
            loss = diff * diff; g = qa > 0.f ? 2.f * diff : 0.f;
            if (g != 0.f) atomicAdd(gb4 + a, g);
        }
        s_g[tid] = g; s_a[tid] = a;
        if (i < S) g_out[i] = g;
        for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
        if (tid == 0 && loss != 0.f) atomicAdd(scalars, loss);
    }
    __syncthreads();
    const float g = s_g[ii]; const int a = s_a[ii];
#pragma unroll 13
    for (int j = jj; j < DQ_K4; j += 8) {
        const float h = (i < n && j < DQ_H3) ? __bfloat162float(h3t[(size_t)j * S + i]) : 0.f;
        const bool on = g != 0.f && h > 0.f;
        const __nv_bfloat16 db = __float2bfloat16_rn(on ? g * __ldg(w4 + (size_t)a * DQ_H3 + j) : 0.f);
        if (i < S) d3t[(size_t)j * S + i] = db;
        tile[ii][j] = db;
    }
    __syncthreads();
    for (int r = 0; r < D3_RAYS; ++r) if (tid < DQ_K4 && i0 + r < S) d3[(size_t)(i0 + r) * DQ_K4 + tid] = tile[r][tid];
}
// dW4[a, :] += g h3 (rank-1 per ray, scatter-added): off the step's critical path -- it runs beside the data path and is only needed when the
// gradients are collected. g_in: the per-ray output-layer gradient k_delta3 left behind.
__global__ void __launch_bounds__(256) k_dw4_rank1(const float* __restrict__ g_in, const uint32_t* __restrict__ actions, int n, int S, const __nv_bfloat16* __restrict__ h3t, float* __restrict__ gw4) {
    const int i = blockIdx.x * D3_RAYS + (threadIdx.x & 31), jj = threadIdx.x >> 5;
    if (i >= n) return;
    const float g = g_in[i];
    if (g == 0.f) return;
    const int a = (int)actions[i];
    const unsigned short* h3 = reinterpret_cast<const unsigned short*>(h3t) + i;
    // four consecutive units per 16-byte vector reduction (a row of W4 is 800 bytes: 16-byte aligned); units behind a dead ReLU add zero
#pragma unroll 7
    for (int j0 = 4 * jj; j0 < DQ_H3; j0 += 32) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float h = __uint_as_float((uint32_t)__ldg(h3 + (size_t)(j0 + e) * S) << 16); v[e] = h > 0.f ? g * h : 0.f; }
        if (v[0] != 0.f || v[1] != 0.f || v[2] != 0.f || v[3] != 0.f)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gw4 + (size_t)a * DQ_H3 + j0), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
}
// Supervised variant (NN_Q_Value_Trainer/Source/main.cu:110-117: loss = sum_batches squared_distance(targets, Q(s)) over ALL
// 144 outputs): the output-layer gradient is dense. g[i][a] = 2 (q_a - y_a) relu'(q_a); one thread per (ray, action).
__global__ void k_g4_full(const float* __restrict__ q, const float* __restrict__ targets, int n, int S, float* __restrict__ g4, float* __restrict__ scalars) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    float loss = 0.f, g = 0.f;
    if (i < n) { const float qa = q[(size_t)a * S + i], diff = qa - targets[(size_t)i * DQ_OUT + a]; loss = diff * diff; g = qa > 0.f ? 2.f * diff : 0.f; }
    if (i < S) g4[(size_t)a * S + i] = g;
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if ((threadIdx.x & 31) == 0 && loss != 0.f) atomicAdd(scalars, loss);
}
// delta3[i][j] = relu'(h3[i][j]) sum_a g[i][a] W4[a][j]; one thread per (ray, hidden unit)
__global__ void k_delta3_full(const float* __restrict__ g4, const float* __restrict__ w4, const __nv_bfloat16* __restrict__ h3t, int n, int S,
                              __nv_bfloat16* __restrict__ d3, __nv_bfloat16* __restrict__ d3t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= S) return;
    float d = 0.f;
    if (i < n && j < DQ_H3 && __bfloat162float(h3t[(size_t)j * S + i]) > 0.f)
        for (int a = 0; a < DQ_OUT; ++a) d += g4[(size_t)a * S + i] * __ldg(w4 + (size_t)a * DQ_H3 + j);
    const __nv_bfloat16 db = __float2bfloat16_rn(d);
    d3[(size_t)i * DQ_K4 + j] = db; d3t[(size_t)j * S + i] = db;
}
// dW4[a][j] = sum_i g[i][a] h3[i][j], db4[a] = sum_i g[i][a]; one warp per (a, j), j == DQ_H3 is the bias
__global__ void k_dw4_full(const float* __restrict__ g4, const __nv_bfloat16* __restrict__ h3t, int n, int S, float* __restrict__ gw4, float* __restrict__ gb4) {
    const int a = blockIdx.y, j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j > DQ_H3) return;
    float acc = 0.f;
    for (int i = lane; i < n; i += 32) acc += g4[(size_t)a * S + i] * (j < DQ_H3 ? __bfloat162float(h3t[(size_t)j * S + i]) : 1.f);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) { if (j < DQ_H3) gw4[(size_t)a * DQ_H3 + j] = acc; else gb4[a] = acc; }
}
// hidden-layer delta: d = pre * relu'(h); written feature-major (weight-gradient GEMM's A) and, for the layer that has one below it, ray-major (next data
// GEMM's A). One CTA = 32 rays x 64 features: all global reads and the feature-major store walk the ray index (coalesced); the ray-major copy goes
// through a shared-memory tile and leaves as 128-byte rows. (One thread per element with a 2-byte store at stride k_pad was 14 us of the step.)
constexpr int DH_RAYS = 32, DH_FEATS = 64;
__global__ void __launch_bounds__(256) k_delta_hidden(const float* __restrict__ pre, int ld_pre, const __nv_bfloat16* __restrict__ ht, int S, int n_feat, int k_pad,
                                                      __nv_bfloat16* __restrict__ d_ray, __nv_bfloat16* __restrict__ d_feat) {
    __shared__ __nv_bfloat16 tile[DH_RAYS][DH_FEATS + 2];                // rows 33 words apart: conflict-free both ways
    const int i0 = blockIdx.x * DH_RAYS, f0 = blockIdx.y * DH_FEATS, lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = i0 + lane;
    RLPT_PDL_SYNC();
#pragma unroll
    for (int q = 0; q < DH_FEATS / 8; ++q) {
        const int fl = w + 8 * q, j = f0 + fl;
        float d = 0.f;
        if (i < S && j < n_feat) { const float h = __bfloat162float(ht[(size_t)j * S + i]); d = h > 0.f ? pre[(size_t)j * ld_pre + i] : 0.f; }       // pre is feature-major (the GEMM stores it transposed)
        const __nv_bfloat16 db = __float2bfloat16_rn(d);
        if (i < S && j < k_pad) d_feat[(size_t)j * S + i] = db;
        tile[lane][fl] = db;
    }
    if (!d_ray) return;
    __syncthreads();
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(&tile[0][0]);
#pragma unroll
    for (int q = 0; q < DH_RAYS / 8; ++q) {
        const int r = w + 8 * q;                                           // ray of the tile; lane = pair of features
        if (i0 + r < S && f0 + 2 * lane < k_pad)
            reinterpret_cast<uint32_t*>(d_ray + (size_t)(i0 + r) * k_pad + f0)[lane] = tw[r * ((DH_FEATS + 2) / 2) + lane];
    }
}
// gather the gradient of every parameter from the GEMM outputs; W1 through the rank-3 identity dW1[j][i] = db1_j v_i - G[j][i % 3]
__global__ void k_collect_grads(const float* __restrict__ dw3x, const float* __restrict__ dw2x, const float* __restrict__ dg, const float* __restrict__ v, int k_in,
                                float* gw1, float* gb1, float* gw2, float* gb2, float* gw3, float* gb3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = DQ_H1 * k_in, n2 = DQ_H2 * DQ_H1, n3 = DQ_H3 * DQ_H2;
    if (i < n1) { int j = i / k_in, c = i % k_in; gw1[i] = dg[j * 16 + 3] * v[c] - dg[j * 16 + c % 3]; return; }
    i -= n1;
    if (i < n2) { int r = i / DQ_H1, c = i % DQ_H1; gw2[i] = dw2x[(size_t)r * DQ_K2 + c]; return; }
    i -= n2;
    if (i < n3) { int r = i / DQ_H2, c = i % DQ_H2; gw3[i] = dw3x[(size_t)r * DQ_K3 + c]; return; }
    i -= n3;
    if (i < DQ_H1) { gb1[i] = dg[i * 16 + 3]; return; }
    i -= DQ_H1;
    if (i < DQ_H2) { gb2[i] = dw2x[(size_t)i * DQ_K2 + DQ_H1]; return; }
    i -= DQ_H2;
    if (i < DQ_H3) gb3[i] = dw3x[(size_t)i * DQ_K3 + DQ_H2];
}
// Adam as DyNet's AdamTrainer applies it: gradient scaled by min(1, clip / ||g||), m and v updated, step size
// lr sqrt(1 - beta2^t) / (1 - beta1^t), x -= step * m / (sqrt(v) + eps)
// The step counter lives on the device (scalars[3], advanced once per step by k_step_begin; k_adam_fused derives the bias-corrected
// step size into scalars[2]), so that a captured optimiser step can be replayed as a CUDA graph without any host value in it.
constexpr int SQN_BLOCKS = 128;        // k_sqnorm_all's / k_collect_norm's grid: that many partial sums, added in a fixed order by k_adam_fused
// The eight parameter arrays (W1 b1 .. W4 b4) are separate allocations; the per-step element-wise passes run over all of them
// in ONE launch each (segment table passed by value) instead of eight: the optimiser step is a chain of ~40 tiny kernels and
// every node costs a few microseconds of dependency latency.
struct ParamSegs { float* x[8]; float* g[8]; float* m[8]; float* v[8]; int start[9]; };
__device__ __forceinline__ int seg_of(const ParamSegs& t, int i) { int s = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) s += i >= t.start[k] ? 1 : 0;
    return s; }
// Start of an optimiser step (runs beside the forward pass): gradients, GEMM outputs and the step's scalars zeroed, Adam's step count advanced, the
// batch inputs of the layer-1 gradient GEMM xt = rows (x, y, z, 1), and zero activations for the padding rays of a partial batch (the forward pass
// writes valid rays only; its "ones" units take care of the bias rows).
__global__ void k_step_begin(ParamSegs t, float* a, int na, float* b, int nb, float* c, int nc, float* scalars, int advance,
                             const float4* __restrict__ pos, int n, int S, __nv_bfloat16* __restrict__ xt, __nv_bfloat16* __restrict__ h1t, __nv_bfloat16* __restrict__ h2t, __nv_bfloat16* __restrict__ h3t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, total = t.start[8];
    if (i < S) {
        const bool ok = i < n; const float4 x = ok ? pos[i] : make_float4(0, 0, 0, 0);
        const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
        xt[i] = __float2bfloat16_rn(x.x); xt[(size_t)S + i] = __float2bfloat16_rn(x.y); xt[(size_t)2 * S + i] = __float2bfloat16_rn(x.z); xt[(size_t)3 * S + i] = __float2bfloat16_rn(ok ? 1.f : 0.f);
        for (int r = 4; r < 16; ++r) xt[(size_t)r * S + i] = zero;
        if (!ok) {
            for (int r = 0; r < DQ_K2; ++r) { h1t[(size_t)r * S + i] = zero; h3t[(size_t)r * S + i] = zero; }
            for (int r = 0; r < DQ_K3; ++r) h2t[(size_t)r * S + i] = zero;
        }
    }
    if (i < total) { const int s = seg_of(t, i); t.g[s][i - t.start[s]] = 0.f; return; }
    int j = i - total;
    if (j < na) { a[j] = 0.f; return; } j -= na;
    if (j < nb) { b[j] = 0.f; return; } j -= nb;
    if (j < nc) { c[j] = 0.f; return; } j -= nc;
    if (j < 2) scalars[j] = 0.f;                               // loss and gradient norm; [2] step size persists
    if (j == 2 && advance) scalars[3] += 1.f;                  // Adam's step count t of THIS step (nothing else writes it)
}
// squared gradient norm, in a FIXED summation order (per-thread strided sums -> block tree -> 128 partials added by one thread): after a gradient
// all-reduce every rank holds the same gradients, and the clipping factor derived from this norm must then be the same bits on every rank, or the
// replicas of the network drift apart in the last place. (A float atomicAdd per warp, the first form, sums in arrival order.)
__global__ void __launch_bounds__(256) k_sqnorm_all(ParamSegs t, float* __restrict__ partial) {
    __shared__ float s_w[8];
    float acc = 0.f; const int total = t.start[8];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) { const int s = seg_of(t, i); const float g = t.g[s][i - t.start[s]]; acc += g * g; }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float b = 0.f; for (int w = 0; w < 8; ++w) b += s_w[w]; partial[blockIdx.x] = b; }
}
// k_collect_grads + k_sqnorm_all in one pass (single-GPU step; with a gradient all-reduce between them the two stay separate): every thread derives
// the gradients it is responsible for from the GEMM outputs (W4 / b4 are already in place, accumulated by k_delta3), stores them and sums their
// squares in k_sqnorm_all's order, so the norm has the same bits either way.
__global__ void __launch_bounds__(1024) k_collect_norm(ParamSegs t, const float* __restrict__ dw3x, const float* __restrict__ dw2x, const float* __restrict__ dg, const float* __restrict__ v, int k_in,
                                                      float* __restrict__ partial) {
    __shared__ float s_w[32];
    float acc = 0.f; const int total = t.start[8];
    RLPT_PDL_SYNC();
#pragma unroll 2
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int s = seg_of(t, i), j = i - t.start[s];
        float g;
        switch (s) {
            case 0: { const int r = j / k_in, c = j - r * k_in; g = dg[r * 16 + 3] * v[c] - dg[r * 16 + c % 3]; break; }
            case 1: g = dg[j * 16 + 3]; break;
            case 2: { const int r = j / DQ_H1, c = j - r * DQ_H1; g = dw2x[(size_t)r * DQ_K2 + c]; break; }
            case 3: g = dw2x[(size_t)j * DQ_K2 + DQ_H1]; break;
            case 4: { const int r = j / DQ_H2, c = j - r * DQ_H2; g = dw3x[(size_t)r * DQ_K3 + c]; break; }
            case 5: g = dw3x[(size_t)j * DQ_K3 + DQ_H2]; break;
            default: g = t.g[s][j]; break;
        }
        if (s < 6) t.g[s][j] = g;
        acc += g * g;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float b = 0.f; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) b += s_w[w]; partial[blockIdx.x] = b; }
}
// k_adam_tick + k_adam_all + k_layer1_operands + k_pack_all in one launch. Every CTA derives the clipping factor (the 128 partial sums, added in
// k_adam_tick's order by its first warp) and the bias-corrected step size (step count t = scalars[3], advanced by k_step_begin) for itself.
// CTAs [0, 200): one CTA per row of W1 -- Adam on the row and on b1, then the row's layer-1 operands c1 = b1 + W1 v, M1 = column sums by coordinate.
// The other CTAs: Adam element-wise on W2 b2 W3 b3 W4 b4, each updated weight written straight into its bf16 operand copies (the forward pass's
// packed B operands, and the transposes the backward data path multiplies by; their padding stays zero from the first pack).
constexpr int ADAM_ROW_BLOCKS = DQ_H1, ADAM_ROW_ITEMS = 4;                // one CTA per row of W1; the first 4 x 256 inputs of the row are requested before the norm is known
__global__ void __launch_bounds__(256) k_adam_fused(ParamSegs t, float* __restrict__ scalars, const float* __restrict__ sq_partial, float* __restrict__ loss_total,
                                                    float lr, float clip, float beta1, float beta2, float eps, const float* __restrict__ v, int k_in, float* __restrict__ c1, float* __restrict__ m1,
                                                    __nv_bfloat16* __restrict__ w2p, __nv_bfloat16* __restrict__ w3p, __nv_bfloat16* __restrict__ w4p, __nv_bfloat16* __restrict__ w3t, __nv_bfloat16* __restrict__ w2t,
                                                    __nv_bfloat16* __restrict__ w3tp, __nv_bfloat16* __restrict__ w2tp) {
    __shared__ float s_n2; __shared__ float s_red[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool rows = blockIdx.x < ADAM_ROW_BLOCKS;
    RLPT_PDL_SYNC();
    // this thread's elements: their gradient and moments are requested before the CTA turns to the norm (everything here is latency)
    int seg[ADAM_ROW_ITEMS], idx[ADAM_ROW_ITEMS]; float g[ADAM_ROW_ITEMS], mm[ADAM_ROW_ITEMS], vv[ADAM_ROW_ITEMS], xx[ADAM_ROW_ITEMS], vin[ADAM_ROW_ITEMS];
    int n_items = 0;
    if (rows) {
#pragma unroll
        for (int q = 0; q < ADAM_ROW_ITEMS; ++q) { const int i = threadIdx.x + 256 * q; if (i < k_in) { seg[q] = 0; idx[q] = blockIdx.x * k_in + i; vin[q] = v[i]; n_items = q + 1; } else { seg[q] = -1; idx[q] = 0; vin[q] = 0.f; } }
    } else {
        const int i = t.start[2] + (blockIdx.x - ADAM_ROW_BLOCKS) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int q = 0; q < ADAM_ROW_ITEMS; ++q) { seg[q] = -1; idx[q] = 0; vin[q] = 0.f; }
        if (i < t.start[8]) { seg[0] = seg_of(t, i); idx[0] = i - t.start[seg[0]]; n_items = 1; }
    }
#pragma unroll
    for (int q = 0; q < ADAM_ROW_ITEMS; ++q) if (seg[q] >= 0) { g[q] = t.g[seg[q]][idx[q]]; mm[q] = t.m[seg[q]][idx[q]]; vv[q] = t.v[seg[q]][idx[q]]; xx[q] = t.x[seg[q]][idx[q]]; }
    float gb = 0.f, mb = 0.f, vb = 0.f, xb1 = 0.f;                          // the row's bias b1 (thread 0 of a row CTA)
    if (rows && threadIdx.x == 0) { gb = t.g[1][blockIdx.x]; mb = t.m[1][blockIdx.x]; vb = t.v[1][blockIdx.x]; xb1 = t.x[1][blockIdx.x]; }
    if (warp == 0) {
        float n2 = (sq_partial[4 * lane] + sq_partial[4 * lane + 1]) + (sq_partial[4 * lane + 2] + sq_partial[4 * lane + 3]);
        for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        if (lane == 0) s_n2 = n2;
    }
    __syncthreads();
    const float n2 = s_n2, tt = scalars[3];
    const float step_size = lr * sqrtf(1.f - powf(beta2, tt)) / (1.f - powf(beta1, tt));
    const float norm = sqrtf(n2), scale = (clip > 0.f && norm > clip) ? clip / norm : 1.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) { scalars[1] = n2; scalars[2] = step_size; if (loss_total) *loss_total += scalars[0]; }
    auto adam = [&](float gi, float& mi, float& vi, float x) {
        gi *= scale;
        mi = beta1 * mi + (1.f - beta1) * gi; vi = beta2 * vi + (1.f - beta2) * gi * gi;
        return x - step_size * mi / (sqrtf(vi) + eps);
    };
    if (rows) {
        // Adam on the row of W1 and on b1, then the row's layer-1 operands c1 = b1 + W1 v, M1 = column sums by coordinate
        float acc = 0.f, m[3] = { 0.f, 0.f, 0.f };
#pragma unroll
        for (int q = 0; q < ADAM_ROW_ITEMS; ++q) if (seg[q] >= 0) {
            const float w = adam(g[q], mm[q], vv[q], xx[q]);
            t.m[0][idx[q]] = mm[q]; t.v[0][idx[q]] = vv[q]; t.x[0][idx[q]] = w;
            acc += w * vin[q]; const int d = (threadIdx.x + 256 * q) % 3; m[0] += d == 0 ? w : 0.f; m[1] += d == 1 ? w : 0.f; m[2] += d == 2 ? w : 0.f;
        }
        for (int i = threadIdx.x + 256 * ADAM_ROW_ITEMS; i < k_in; i += 256) {          // rows wider than 1024 inputs (large scenes): the rest, plainly
            const int e = blockIdx.x * k_in + i; float mi = t.m[0][e], vi = t.v[0][e];
            const float w = adam(t.g[0][e], mi, vi, t.x[0][e]);
            t.m[0][e] = mi; t.v[0][e] = vi; t.x[0][e] = w;
            acc += w * v[i]; const int d = i % 3; m[0] += d == 0 ? w : 0.f; m[1] += d == 1 ? w : 0.f; m[2] += d == 2 ? w : 0.f;
        }
        for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); for (int d = 0; d < 3; ++d) m[d] += __shfl_xor_sync(0xffffffffu, m[d], o); }
        if (lane == 0) { s_red[warp][0] = acc; s_red[warp][1] = m[0]; s_red[warp][2] = m[1]; s_red[warp][3] = m[2]; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float r4[4] = { 0.f, 0.f, 0.f, 0.f };
            for (int w = 0; w < 8; ++w) for (int d = 0; d < 4; ++d) r4[d] += s_red[w][d];
            const int row = blockIdx.x;
            const float b = adam(gb, mb, vb, xb1);
            t.m[1][row] = mb; t.v[1][row] = vb; t.x[1][row] = b;
            c1[row] = b + r4[0]; m1[3 * row] = r4[1]; m1[3 * row + 1] = r4[2]; m1[3 * row + 2] = r4[3];
        }
        return;
    }
    if (n_items == 0) return;
    // element-wise on W2 b2 W3 b3 W4 b4; each updated weight goes straight into its bf16 operand copies (padding stays zero from the first pack)
    const int s = seg[0], j = idx[0];
    const float x = adam(g[0], mm[0], vv[0], xx[0]);
    t.m[s][j] = mm[0]; t.v[s][j] = vv[0]; t.x[s][j] = x;
    const __nv_bfloat16 xb = __float2bfloat16_rn(x);
    if (s == 2) { const int r = j / DQ_H1, c = j - r * DQ_H1; { const size_t o = wpack_offset(DQ_L2_SPLIT, r, c, DQ_N2, DQ_K2); for (int q = 0; q < DQ_REPLICAS; ++q) *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w2p) + q * DQ_W2P_STRIDE + o) = xb; } if (w2t) { w2t[(size_t)c * DQ_K3 + r] = xb; *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w2tp) + wpack_offset(0, c, r, DQ_N3, DQ_K3)) = xb; } }
    else if (s == 4) { const int r = j / DQ_H2, c = j - r * DQ_H2; { const size_t o = wpack_offset(0, r, c, DQ_N3, DQ_K3); for (int q = 0; q < DQ_REPLICAS; ++q) *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w3p) + q * DQ_W3P_STRIDE + o) = xb; } if (w3t) { w3t[(size_t)c * DQ_K2 + r] = xb; *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w3tp) + wpack_offset(DQ_L2_SPLIT, c, r, DQ_N2, DQ_K2)) = xb; } }
    else if (s == 6) { const int r = j / DQ_H3, c = j - r * DQ_H3; { const size_t o = wpack_offset(0, r, c, DQ_N4, DQ_K4); for (int q = 0; q < DQ_REPLICAS; ++q) *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w4p) + q * DQ_W4P_STRIDE + o) = xb; } }
}
static ParamSegs param_segs(const DqnDev& d, const DqnTrain& t) {
    ParamSegs p{}; DqnHost shape; shape.k_in = d.k_in; int at = 0;
    for (int l = 0; l < 4; ++l) {
        const int nw = DqnHost::rows(l) * shape.cols(l), nb = DqnHost::rows(l);
        p.x[2 * l] = d.w[l]; p.g[2 * l] = t.gw[l]; p.m[2 * l] = t.mw[l]; p.v[2 * l] = t.vw[l]; p.start[2 * l] = at; at += nw;
        p.x[2 * l + 1] = d.b[l]; p.g[2 * l + 1] = t.gb[l]; p.m[2 * l + 1] = t.mb[l]; p.v[2 * l + 1] = t.vb[l]; p.start[2 * l + 1] = at; at += nb;
    }
    p.start[8] = at;
    return p;
}
void dqn_train_free(DqnTrain& t) {
    cudaFree(t.gall); cudaFree(t.sq_partial);
    for (int l = 0; l < 4; ++l) { cudaFree(t.mw[l]); cudaFree(t.mb[l]); cudaFree(t.vw[l]); cudaFree(t.vb[l]); }
    cudaFree(t.dw3x); cudaFree(t.dw2x); cudaFree(t.dg); cudaFree(t.w3t); cudaFree(t.w2t); cudaFree(t.w3tp); cudaFree(t.w2tp); cudaFree(t.h1t); cudaFree(t.h2t); cudaFree(t.h3t); cudaFree(t.xt);
    cudaFree(t.d3); cudaFree(t.d2); cudaFree(t.d3t); cudaFree(t.d2t); cudaFree(t.d1t); cudaFree(t.g3); cudaFree(t.mask1); cudaFree(t.mask2); cudaFree(t.mask3); cudaFree(t.p2); cudaFree(t.p1); cudaFree(t.q); cudaFree(t.scalars); cudaFree(t.g4);
    if (t.side) cudaStreamDestroy(t.side);
    if (t.side2) cudaStreamDestroy(t.side2);
    if (t.side3) cudaStreamDestroy(t.side3);
    for (cudaEvent_t e : t.ev) if (e) cudaEventDestroy(e);
    t = DqnTrain{};
}
int dqn_train_alloc(DqnTrain& t, const DqnDev& d, int capacity) {
    const int S = (capacity + DQ_TILE - 1) / DQ_TILE * DQ_TILE;
    DqnHost shape; shape.k_in = d.k_in;
    const bool fresh = t.gw[0] == nullptr;
    if (fresh) {
        // the eight gradient arrays live in ONE block (W1 b1 .. W4 b4, each 16-byte aligned): a multi-GPU step sums them with a single all-reduce
        size_t gtotal = 0;
        for (int l = 0; l < 4; ++l) gtotal += ((size_t)DqnHost::rows(l) * shape.cols(l) + 3) / 4 * 4 + ((size_t)DqnHost::rows(l) + 3) / 4 * 4;
        DQ_CK(cudaMalloc(&t.gall, 4 * gtotal)); DQ_CK(cudaMemset(t.gall, 0, 4 * gtotal)); t.gall_count = gtotal;
        float* gp = t.gall;
        for (int l = 0; l < 4; ++l) {
            const size_t nw = (size_t)DqnHost::rows(l) * shape.cols(l), nb = DqnHost::rows(l);
            t.gw[l] = gp; gp += (nw + 3) / 4 * 4; t.gb[l] = gp; gp += (nb + 3) / 4 * 4;
            DQ_CK(cudaMalloc(&t.mw[l], 4 * nw)); DQ_CK(cudaMalloc(&t.mb[l], 4 * nb));
            DQ_CK(cudaMalloc(&t.vw[l], 4 * nw)); DQ_CK(cudaMalloc(&t.vb[l], 4 * nb));
            DQ_CK(cudaMemset(t.mw[l], 0, 4 * nw)); DQ_CK(cudaMemset(t.mb[l], 0, 4 * nb)); DQ_CK(cudaMemset(t.vw[l], 0, 4 * nw)); DQ_CK(cudaMemset(t.vb[l], 0, 4 * nb));
        }
        DQ_CK(cudaMalloc(&t.dw3x, 4 * (size_t)DQ_N3 * DQ_K3)); DQ_CK(cudaMalloc(&t.dw2x, 4 * (size_t)DQ_N2 * DQ_K2)); DQ_CK(cudaMalloc(&t.dg, 4 * (size_t)DQ_K2 * 16));
        DQ_CK(cudaMalloc(&t.w3t, 2 * (size_t)DQ_N2 * DQ_K2)); DQ_CK(cudaMalloc(&t.w2t, 2 * (size_t)DQ_N3 * DQ_K3)); DQ_CK(cudaMalloc(&t.w3tp, 2 * (size_t)DQ_N2 * DQ_K2)); DQ_CK(cudaMalloc(&t.w2tp, 2 * (size_t)DQ_N3 * DQ_K3)); DQ_CK(cudaMalloc(&t.scalars, 4 * 4)); DQ_CK(cudaMemset(t.scalars, 0, 4 * 4)); DQ_CK(cudaMalloc(&t.sq_partial, 4 * 128));
        t.step = 0;
    }
    if (S > t.capacity) {
        cudaFree(t.h1t); cudaFree(t.h2t); cudaFree(t.h3t); cudaFree(t.xt); cudaFree(t.d3); cudaFree(t.d2); cudaFree(t.d3t); cudaFree(t.d2t); cudaFree(t.d1t); cudaFree(t.g3); cudaFree(t.mask1); cudaFree(t.mask2); cudaFree(t.mask3); cudaFree(t.p2); cudaFree(t.p1); cudaFree(t.q);
        DQ_CK(cudaMalloc(&t.h1t, 2 * (size_t)DQ_K2 * S)); DQ_CK(cudaMalloc(&t.h2t, 2 * (size_t)DQ_K3 * S)); DQ_CK(cudaMalloc(&t.h3t, 2 * (size_t)DQ_K4 * S)); DQ_CK(cudaMalloc(&t.xt, 2 * (size_t)16 * S));
        DQ_CK(cudaMalloc(&t.d3, 2 * (size_t)S * DQ_K4)); DQ_CK(cudaMalloc(&t.d2, 2 * (size_t)S * DQ_K3)); DQ_CK(cudaMalloc(&t.d3t, 2 * (size_t)DQ_K4 * S)); DQ_CK(cudaMalloc(&t.d2t, 2 * (size_t)DQ_K3 * S));
        DQ_CK(cudaMalloc(&t.d1t, 2 * (size_t)DQ_K2 * S)); DQ_CK(cudaMalloc(&t.g3, 4 * (size_t)S));
        DQ_CK(cudaMalloc(&t.mask1, 4 * (size_t)((DQ_K2 + 31) / 32) * S)); DQ_CK(cudaMalloc(&t.mask2, 4 * (size_t)((DQ_K3 + 31) / 32) * S)); DQ_CK(cudaMalloc(&t.mask3, 4 * (size_t)((DQ_K4 + 31) / 32) * S));
        DQ_CK(cudaMemset(t.mask1, 0, 4 * (size_t)((DQ_K2 + 31) / 32) * S)); DQ_CK(cudaMemset(t.mask2, 0, 4 * (size_t)((DQ_K3 + 31) / 32) * S)); DQ_CK(cudaMemset(t.mask3, 0, 4 * (size_t)((DQ_K4 + 31) / 32) * S)); DQ_CK(cudaMalloc(&t.p2, 4 * (size_t)S * DQ_N2)); DQ_CK(cudaMalloc(&t.p1, 4 * (size_t)S * DQ_N3)); DQ_CK(cudaMalloc(&t.q, 4 * (size_t)DQ_OUT * S));
        t.capacity = S;
    }
    return 0;
}
// out = packed (wpack_offset) bf16 copy of W^T for k_dqn_backward: W is [k][n] row-major, the operand has n_pad rows (W's columns) x k_pad inputs (W's rows)
__global__ void k_pack_transposed(const float* __restrict__ w, int n, int k, int n_pad, int k_pad, int n_split, __nv_bfloat16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * k_pad) return;
    const int row = i / k_pad, col = i % k_pad;
    const float v = (row < n && col < k) ? w[(size_t)col * n + row] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(out) + wpack_offset(n_split, row, col, n_pad, k_pad)) = __float2bfloat16_rn(v);
}
static int refresh_transposes(const DqnDev& d, DqnTrain& t, cudaStream_t s) {
    k_pack_transposed<<<(DQ_N2 * DQ_K2 + 255) / 256, 256, 0, s>>>(d.w[2], DQ_H2, DQ_H3, DQ_N2, DQ_K2, DQ_L2_SPLIT, t.w3tp);       // W3 [200][300] -> W3^T as [304][208]
    k_pack_transposed<<<(DQ_N3 * DQ_K3 + 255) / 256, 256, 0, s>>>(d.w[1], DQ_H1, DQ_H2, DQ_N3, DQ_K3, 0, t.w2tp);                 // W2 [300][200] -> W2^T as [208][304]
    // W3 is [200][300]: W3^T as [304][208] (row = layer-2 unit, K = layer-3 unit); W2 is [300][200]: W2^T as [208][304]
    k_transpose_bf16<<<(DQ_N2 * DQ_K2 + 255) / 256, 256, 0, s>>>(d.w[2], DQ_H3, DQ_H2, DQ_N2, DQ_K2, t.w3t);
    k_transpose_bf16<<<(DQ_N3 * DQ_K3 + 255) / 256, 256, 0, s>>>(d.w[1], DQ_H2, DQ_H1, DQ_N3, DQ_K3, t.w2t);
    return (int)cudaGetLastError();
}

DqnFwdParams dqn_train_forward_params(const DqnDev& d, DqnTrain& t, const float4* pos, int n) {
    const int S = (n + DQ_TILE - 1) / DQ_TILE * DQ_TILE;
    DqnFwdParams fp{}; fp.pos = pos; fp.n = n; fp.c1 = d.c1; fp.m1 = d.m1; fp.b2 = d.b[1]; fp.b3 = d.b[2]; fp.b4 = d.b[3]; fp.w2p = d.w2p; fp.w3p = d.w3p; fp.w4p = d.w4p;
    fp.q = t.q; fp.q_stride = S; fp.h1t = t.h1t; fp.h2t = t.h2t; fp.h3t = t.h3t; fp.h_stride = S; fp.mask1 = t.mask1; fp.mask2 = t.mask2; fp.mask3 = t.mask3;
    return fp;
}
int dqn_train_prepare(DqnDev& d, DqnTrain& t, int n, cudaStream_t s) {           // everything a captured step must not contain: allocations, the first transposes, the side stream
    int rc = dqn_train_alloc(t, d, n); if (rc) return rc;
    if (!t.transposes_fresh) { rc = refresh_transposes(d, t, s); if (rc) return rc; t.transposes_fresh = true; }
    if (!t.side) {
        if (const char* e = getenv("RLPT_NQ_PDL")) t.pdl = atoi(e) != 0;
        if (const char* e = getenv("RLPT_NQ_FUSED_BWD")) t.fused_bwd = atoi(e) != 0;
        DQ_CK(cudaStreamCreateWithFlags(&t.side, cudaStreamNonBlocking)); DQ_CK(cudaStreamCreateWithFlags(&t.side2, cudaStreamNonBlocking)); DQ_CK(cudaStreamCreateWithFlags(&t.side3, cudaStreamNonBlocking));
        for (cudaEvent_t& e : t.ev) DQ_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return 0;
}
// Start of a step, beside whatever the caller enqueues next on `s` (the forward pass): zeroing and the batch inputs on the side stream.
int dqn_train_begin(DqnDev& d, DqnTrain& t, const float4* pos, int n, bool advance, cudaStream_t s) {
    if (!d.ready || n <= 0) return n == 0 ? 0 : -1;
    int rc = dqn_train_prepare(d, t, n, s); if (rc) return rc;               // (afterwards k_adam_fused keeps operands and transposes current)
    const int S = (n + DQ_TILE - 1) / DQ_TILE * DQ_TILE;
    const ParamSegs segs = param_segs(d, t);
    const int na = DQ_N3 * DQ_K3, nb = DQ_N2 * DQ_K2, nc = DQ_K2 * 16; int total = segs.start[8] + na + nb + nc + 3; if (total < S) total = S;
    DQ_CK(cudaEventRecord(t.ev[3], s)); DQ_CK(cudaStreamWaitEvent(t.side, t.ev[3], 0));
    k_step_begin<<<(total + 255) / 256, 256, 0, t.side>>>(segs, t.dw3x, na, t.dw2x, nb, t.dg, nc, t.scalars, advance ? 1 : 0, pos, n, S, t.xt, t.h1t, t.h2t, t.h3t);
    DQ_CK(cudaEventRecord(t.ev[4], t.side));
    t.begun = true;
    return (int)cudaGetLastError();
}
int dqn_train_batch(DqnDev& d, DqnTrain& t, const float4* pos, const uint32_t* actions, const float* targets, int n, bool apply_update,
                    dqn_allreduce_fn allreduce, void* allreduce_user, cudaStream_t s, bool all_outputs, float* loss_total, bool forward_done, const DqnTdParams* tdp) {
    if (!d.ready || n <= 0) return n == 0 ? 0 : -1;
    int rc = 0;
    if (!t.begun) { rc = dqn_train_begin(d, t, pos, n, apply_update, s); if (rc) return rc; }
    t.begun = false;
    const int S = (n + DQ_TILE - 1) / DQ_TILE * DQ_TILE;
    const ParamSegs segs = param_segs(d, t);
    const bool pdl = t.pdl && !all_outputs;
    // forward, activations kept
    if (!forward_done) { const DqnFwdParams fp = dqn_train_forward_params(d, t, pos, n); rc = dqn_forward(d, fp, s); if (rc) return rc; }
    DQ_CK(cudaStreamWaitEvent(s, t.ev[4], 0));                              // the zeroing and the batch inputs are in place
    // backward: data path
    if (all_outputs) {            // targets: [n][144]
        if (S > t.g4_capacity) { cudaFree(t.g4); t.g4 = nullptr; t.g4_capacity = 0; DQ_CK(cudaMalloc(&t.g4, 4 * (size_t)DQ_OUT * S)); t.g4_capacity = S; }
        k_g4_full<<<dim3((S + 127) / 128, DQ_OUT), 128, 0, s>>>(t.q, targets, n, S, t.g4, t.scalars);
        k_delta3_full<<<dim3((S + 127) / 128, DQ_K4), 128, 0, s>>>(t.g4, d.w[3], t.h3t, n, S, t.d3, t.d3t);
        k_dw4_full<<<dim3((DQ_H3 + 1 + 7) / 8, DQ_OUT), 256, 0, s>>>(t.g4, t.h3t, n, S, t.gw[3], t.gb[3]);
    } else if (t.fused_bwd) {
        // TD step: output-layer delta, both data products and both masks in ONE kernel; the three weight-gradient GEMMs then run side by side
        DqnBwdParams bp{}; bp.n = n; bp.S = S; bp.q = t.q; bp.actions = actions; bp.targets = const_cast<float*>(targets); if (tdp) bp.td = *tdp;
        bp.w4 = d.w[3]; bp.mask1 = t.mask1; bp.mask2 = t.mask2; bp.mask3 = t.mask3; bp.w3tp = t.w3tp; bp.w2tp = t.w2tp; bp.d3t = t.d3t; bp.d2t = t.d2t; bp.d1t = t.d1t;
        bp.g_out = t.g3; bp.gb4 = t.gb[3]; bp.scalars = t.scalars;
        rc = dqn_backward(bp, s); if (rc) return rc;
        const int ks = S >= 2048 ? 16 : (S >= 512 ? 4 : 1);
        DQ_CK(cudaEventRecord(t.ev[0], s)); DQ_CK(cudaStreamWaitEvent(t.side, t.ev[0], 0)); DQ_CK(cudaStreamWaitEvent(t.side2, t.ev[0], 0));
        rc = gemm_tn(t.d3t, S, t.h2t, S, t.dw3x, DQ_K3, DQ_K4, DQ_K3, S, ks, t.side); if (rc) return rc;                 // [208 x S] x [304 x S]^T
        rc = gemm_tn(t.d2t, S, t.h1t, S, t.dw2x, DQ_K2, DQ_K3, DQ_K2, S, ks, t.side2); if (rc) return rc;                // [304 x S] x [208 x S]^T
        rc = gemm_tn(t.d1t, S, t.xt, S, t.dg, 16, DQ_K2, 16, S, ks, s); if (rc) return rc;                                // [208 x S] x [16 x S]^T
        k_dw4_rank1<<<(S + D3_RAYS - 1) / D3_RAYS, 256, 0, s>>>(t.g3, actions, n, S, t.h3t, t.gw[3]);       // (on a stream of its own: 85 -> 88 us per step)
        DQ_CK(cudaEventRecord(t.ev[2], t.side)); DQ_CK(cudaStreamWaitEvent(s, t.ev[2], 0));
        DQ_CK(cudaEventRecord(t.ev[1], t.side2)); DQ_CK(cudaStreamWaitEvent(s, t.ev[1], 0));
    }
    if (all_outputs || !t.fused_bwd) {
    if (!all_outputs) {
        DqnTdParams td{}; if (tdp) td = *tdp;
        DQ_CK(launch_ex(pdl, k_delta3, dim3((S + D3_RAYS - 1) / D3_RAYS), dim3(256), 0, s, t.q, actions, const_cast<float*>(targets), td, n, S, d.w[3], t.h3t, t.d3, t.d3t, t.g3, t.gb[3], t.scalars));
    }
    // The weight-gradient products need only the deltas of their own layer, not the rest of the data path: they run on a side stream
    // (forked and joined with events, which also works inside a stream capture: the captured graph gets the parallel branches), so
    // the chain the step waits for is  d3 -> p2 -> d2 -> p1 -> d1 -> dG  with dW3, dW2 beside it.
    const int ks = S >= 2048 ? 16 : (S >= 512 ? 4 : 1);
    static const int bn_env = getenv("RLPT_GM_BN") ? atoi(getenv("RLPT_GM_BN")) : 80;     // output columns per CTA of the two data-path GEMMs (80: 128 CTAs at batch 4096)
    const int bn_data = (bn_env >= 16 && bn_env <= GM_BN && bn_env % 16 == 0) ? bn_env : GM_BN;
    DQ_CK(cudaEventRecord(t.ev[0], s)); DQ_CK(cudaStreamWaitEvent(t.side, t.ev[0], 0));
    if (!all_outputs) k_dw4_rank1<<<(S + D3_RAYS - 1) / D3_RAYS, 256, 0, t.side>>>(t.g3, actions, n, S, t.h3t, t.gw[3]);
    rc = gemm_tn(t.d3t, S, t.h2t, S, t.dw3x, DQ_K3, DQ_K4, DQ_K3, S, ks, t.side); if (rc) return rc;                 // [208 x S] x [304 x S]^T
    rc = gemm_tn(t.d3, DQ_K4, t.w3t, DQ_K2, t.p2, S, S, DQ_N2, DQ_K4, 1, s, 1, pdl, bn_data); if (rc) return rc;                    // [S x 208] x [304 x 208]^T, stored [304][S]
    // (the mask + bf16 conversion stays a launch of its own: fused into the GEMM's epilogue it ran on the GEMM's 64 CTAs, 141 -> 191 us per step)
    DQ_CK(launch_ex(pdl, k_delta_hidden, dim3((S + DH_RAYS - 1) / DH_RAYS, (DQ_K3 + DH_FEATS - 1) / DH_FEATS), dim3(256), 0, s, t.p2, S, t.h2t, S, DQ_H2, DQ_K3, t.d2, t.d2t));
    DQ_CK(cudaEventRecord(t.ev[1], s)); DQ_CK(cudaStreamWaitEvent(t.side, t.ev[1], 0));
    rc = gemm_tn(t.d2t, S, t.h1t, S, t.dw2x, DQ_K2, DQ_K3, DQ_K2, S, ks, t.side); if (rc) return rc;                 // [304 x S] x [208 x S]^T
    rc = gemm_tn(t.d2, DQ_K3, t.w2t, DQ_K3, t.p1, S, S, DQ_N3, DQ_K3, 1, s, 1, pdl, bn_data); if (rc) return rc;                    // [S x 304] x [208 x 304]^T, stored [208][S]
    DQ_CK(launch_ex(pdl, k_delta_hidden, dim3((S + DH_RAYS - 1) / DH_RAYS, (DQ_K2 + DH_FEATS - 1) / DH_FEATS), dim3(256), 0, s, t.p1, S, t.h1t, S, DQ_H1, DQ_K2, (__nv_bfloat16*)nullptr, t.d1t));
    rc = gemm_tn(t.d1t, S, t.xt, S, t.dg, 16, DQ_K2, 16, S, ks, s, 0, pdl); if (rc) return rc;                                // [208 x S] x [16 x S]^T
    DQ_CK(cudaEventRecord(t.ev[2], t.side)); DQ_CK(cudaStreamWaitEvent(s, t.ev[2], 0));
    }
    if (allreduce) {
        const int n_collect = DQ_H1 * d.k_in + DQ_H2 * DQ_H1 + DQ_H3 * DQ_H2 + DQ_H1 + DQ_H2 + DQ_H3;
        k_collect_grads<<<(n_collect + 255) / 256, 256, 0, s>>>(t.dw3x, t.dw2x, t.dg, d.vertices, d.k_in, t.gw[0], t.gb[0], t.gw[1], t.gb[1], t.gw[2], t.gb[2]);
        if (allreduce(t.gall, (uint64_t)t.gall_count, 0, (void*)s, allreduce_user)) return -2;            // all eight gradient arrays at once (padding words are zero)
        if (allreduce(t.scalars, 1, 0, (void*)s, allreduce_user)) return -2;
        if (apply_update) k_sqnorm_all<<<SQN_BLOCKS, 256, 0, s>>>(segs, t.sq_partial);
    } else
        DQ_CK(launch_ex(pdl, k_collect_norm, dim3(SQN_BLOCKS), dim3(1024), 0, s, segs, t.dw3x, t.dw2x, t.dg, d.vertices, d.k_in, t.sq_partial));
    if (!apply_update) return (int)cudaGetLastError();
    t.step++;
    // Adam on all parameters; operands for the next forward / backward refreshed in the same launch (layer-1 rank-3 form, packed bf16 weights, the two transposes)
    {
        const int rest = segs.start[8] - segs.start[2];
        DQ_CK(launch_ex(pdl, k_adam_fused, dim3(ADAM_ROW_BLOCKS + (rest + 255) / 256), dim3(256), 0, s, segs, t.scalars, t.sq_partial, loss_total, t.lr, t.clip, t.beta1, t.beta2, t.eps, d.vertices, d.k_in, d.c1, d.m1,
                        d.w2p, d.w3p, d.w4p, t.w3t, t.w2t, t.w3tp, t.w2tp));
    }
    return (int)cudaGetLastError();
}

int dqn_gemm_set_smem_limit() { return (int)cudaFuncSetAttribute(k_gemm_bf16_tn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_TOTAL); }

}  // namespace rlpt
