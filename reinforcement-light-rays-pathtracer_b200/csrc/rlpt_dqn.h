// rlpt_dqn.h -- the Neural-Q network (DQN) of the reference: K -> 200 -> 300 -> 200 -> 144, ReLU after every layer
// including the output (N/dq_network.cu:8-49, N/fc_layer.cu:29-72), on 5th-generation tensor cores (tcgen05).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace rlpt {

constexpr int DQ_H1 = 200, DQ_H2 = 300, DQ_H3 = 200, DQ_OUT = 144;
// MMA shapes: K padded to a multiple of 16 (bf16 UMMA_K), N to a multiple of 16 (UMMA_N granularity at M = 128)
constexpr int DQ_K2 = 208, DQ_N2 = 304, DQ_K3 = 304, DQ_N3 = 208, DQ_K4 = 208, DQ_N4 = 144;
constexpr int DQ_TILE = 128;        // rays per CTA tile = UMMA_M = TMEM lanes
}  // namespace rlpt
#include "rlpt_dqn_layout.h"
namespace rlpt {
constexpr int DQ_CHUNK = 64;        // weight rows (outputs) staged per shared-memory buffer

// Host copy of the parameters in the reference's order (DyNet TextFileSaver blocks /_0 .. /_7), row-major [out][in].
struct DqnHost {
    int k_in = 0;                                   // input width: 9 floats per triangle (scene vertices) -- 342 Cornell, 918 archway
    std::vector<float> w[4], b[4];                  // w[0]: 200 x k_in, w[1]: 300 x 200, w[2]: 200 x 300, w[3]: 144 x 200
    static int rows(int l) { return l == 0 ? DQ_H1 : l == 1 ? DQ_H2 : l == 2 ? DQ_H3 : DQ_OUT; }
    int cols(int l) const { return l == 0 ? k_in : l == 1 ? DQ_H1 : l == 2 ? DQ_H2 : DQ_H3; }
};

// Device state. fp32 master parameters (the optimiser works on these) and the operands derived from them:
//   layer 1 is affine in the 3-vector x because the input is (v_i - x) for every scene vertex v_i (nn_rendering_helpers.cu:280-298):
//   W1 (v - 1 (x) x) + b1 = c1 - M1 x with c1 = b1 + W1 v, M1[:, d] = sum of the columns i of W1 with i % 3 == d; evaluated in fp32.
//   layers 2..4: bf16 copies in the tcgen05 shared-memory operand layout (K-major, no swizzle, 8x8 core matrices), cut into the chunks the kernel copies.
struct DqnDev {
    int k_in = 0;
    float* w[4] = { nullptr, nullptr, nullptr, nullptr }; float* b[4] = { nullptr, nullptr, nullptr, nullptr };
    float* vertices = nullptr;                      // [k_in] scene vertices, the network's constant input part
    float *c1 = nullptr, *m1 = nullptr;             // [200], [200][3]
    __nv_bfloat16 *w2p = nullptr, *w3p = nullptr, *w4p = nullptr;       // packed B operands as k_dqn_forward streams them (wpack_offset in rlpt_dqn.cu: N parts x K chunks)
    bool ready = false;
};

struct DqnFwdParams {
    const float4* pos; int n;                       // ray positions (xyz)
    const int* n_ptr;                               // when non-null the ray count is read from device memory (wavefront live count)
    const float *c1, *m1, *b2, *b3, *b4;
    const __nv_bfloat16 *w2p, *w3p, *w4p;
    float* q; int q_stride;                         // out: [144][q_stride], action-major (coalesced for producer and consumers)
    __nv_bfloat16 *h1t, *h2t, *h3t; int h_stride;   // optional: activations kept for the backward pass, feature-major [K_pad][h_stride]
    uint32_t *mask1, *mask2, *mask3;                // optional, with the kept activations: relu'(h) as bit words [K_pad / 32][h_stride] (unit j = bit j % 32 of word j / 32)
    const float4* pos2; int n2; float* q2; int q_stride2;       // optional second batch evaluated by the same launch (no activations kept): tiles follow the first batch's
};

// Training state: gradients, Adam moments (DyNet AdamTrainer defaults: lr 1e-3, beta1 0.9, beta2 0.999, eps 1e-8, gradient
// clipping at norm 5), batch activations. Layouts for a batch of n rays, S = n rounded up to 128:
//   h1t [208][S], h2t [304][S], h3t [208][S]  bf16, feature-major (k_dqn_forward keeps them); rows 200 of h1t and 300 of h2t
//   are set to one so that the weight-gradient GEMMs also produce the bias gradients; xt [16][S] = (x, y, z, 1, 0...)
//   d3 [S][208], d2 [S][304] ray-major and d3t [208][S], d2t [304][S], d1t [208][S] feature-major deltas (bf16)
struct DqnTrain {
    int capacity = 0;                               // S the buffers were allocated for
    float *gw[4] = { nullptr, nullptr, nullptr, nullptr }, *gb[4] = { nullptr, nullptr, nullptr, nullptr };      // gradients: views into gall
    float* gall = nullptr; size_t gall_count = 0;   // one block holding all eight gradient arrays
    float* sq_partial = nullptr;                    // [128] per-block partial sums of the squared gradient norm (summed in a fixed order)
    float *mw[4] = { nullptr, nullptr, nullptr, nullptr }, *mb[4] = { nullptr, nullptr, nullptr, nullptr };      // Adam first moments
    float *vw[4] = { nullptr, nullptr, nullptr, nullptr }, *vb[4] = { nullptr, nullptr, nullptr, nullptr };      // Adam second moments
    float *dw3x = nullptr, *dw2x = nullptr, *dg = nullptr;      // GEMM outputs: [208][304] (col 300 = db3), [304][208] (col 200 = db2), [208][16] (cols 0-2 = G, col 3 = db1)
    __nv_bfloat16 *w3t = nullptr, *w2t = nullptr;   // transposed bf16 copies: W3^T [304][208], W2^T [208][304]
    __nv_bfloat16 *w3tp = nullptr, *w2tp = nullptr; // the same transposes as k_dqn_backward streams them (packed chunks, wpack_offset)
    __nv_bfloat16 *h1t = nullptr, *h2t = nullptr, *h3t = nullptr, *xt = nullptr;
    __nv_bfloat16 *d3 = nullptr, *d2 = nullptr, *d3t = nullptr, *d2t = nullptr, *d1t = nullptr;
    uint32_t *mask1 = nullptr, *mask2 = nullptr, *mask3 = nullptr;      // relu'(h1), relu'(h2), relu'(h3) as bit words [K_pad / 32][S] (k_dqn_forward -> k_dqn_backward)
    float* g3 = nullptr;                            // [S] per-ray output-layer gradient of the TD step (k_delta3 -> k_dw4_rank1)
    float *p2 = nullptr, *p1 = nullptr;             // pre-activation deltas [S][304], [S][208] (fp32 GEMM outputs)
    float* q = nullptr;                             // [144][S] predictions of the batch
    float* scalars = nullptr;                       // [0] loss sum, [1] squared gradient norm, [2] Adam step size, [3] Adam step count (kept on the device: graph replay)
    float* g4 = nullptr; int g4_capacity = 0;       // supervised step: dense output-layer gradient [144][S]
    unsigned long long step = 0;
    bool transposes_fresh = false;                  // W3^T / W2^T match the current parameters (k_pack_all refreshes them after every update)
    float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, clip = 5.f;
    cudaStream_t side = nullptr; cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };      // the weight-gradient GEMMs and the step's zeroing run beside the data path
    bool pdl = false;                               // programmatic dependent launch along the step's main chain (RLPT_NQ_PDL=1; measured: 100.9 -> 102.4 us per step, off)
    bool fused_bwd = true;                          // TD step: the backward data path as one kernel (RLPT_NQ_FUSED_BWD=0: delta kernel + GEMMs + mask kernels)
    cudaStream_t side2 = nullptr, side3 = nullptr;
    bool begun = false;                             // dqn_train_begin has been enqueued for the step dqn_train_batch is about to run
};
int dqn_train_alloc(DqnTrain& t, const DqnDev& d, int capacity);
void dqn_train_free(DqnTrain& t);
// TD targets derived inside the output-layer delta kernel (q_next null: targets are read from the array instead)
struct DqnTdParams { const float* q_next; int q_stride; const uint32_t* state; const float* reward; const float* discount; };
void dqn_upload_cell_cos(const float* cos144);
// k_dqn_backward's arguments (rlpt_dqn.cu): the TD step's backward data path of a batch
struct DqnBwdParams {
    int n, S; const float* q; const uint32_t* actions; float* targets; DqnTdParams td; const float* w4;
    const uint32_t *mask1, *mask2, *mask3; const __nv_bfloat16 *w3tp, *w2tp; __nv_bfloat16 *d3t, *d2t, *d1t; float *g_out, *gb4, *scalars;
};
int dqn_backward(const DqnBwdParams& p, cudaStream_t s);
// One optimiser step on a batch (G/deep_learning/neural_q_pathtracer.cu:476-512): forward with kept activations, loss
// sum_b (target_b - Q(s_b)[a_b])^2, backward, Adam update, operands refreshed. pos/actions/targets are device pointers.
// all-reduce hook (may be null): sums the gradient buffers across ranks before the update.
typedef int (*dqn_allreduce_fn)(void* d_buf, uint64_t count, int dtype, void* cuda_stream, void* user);
int dqn_train_batch(DqnDev& d, DqnTrain& t, const float4* pos, const uint32_t* actions, const float* targets, int n, bool apply_update,
                    dqn_allreduce_fn allreduce, void* allreduce_user, cudaStream_t s, bool all_outputs = false,    // all_outputs: targets is [n][144], the supervised loss over every output (actions unused)
                    float* loss_total = nullptr,                                                                   // device scalar the step's loss is added to (may be null)
                    bool forward_done = false,                                                                     // the caller already ran the kept-activation forward into t.q / t.h*t (dqn_train_forward_params)
                    const DqnTdParams* tdp = nullptr);                                                             // TD targets from the next states' Q values, computed in the step (targets: where they are stored)
int dqn_train_begin(DqnDev& d, DqnTrain& t, const float4* pos, int n, bool advance, cudaStream_t s);              // optional: the step's zeroing / batch inputs enqueued beside the caller's own forward launch
DqnFwdParams dqn_train_forward_params(const DqnDev& d, DqnTrain& t, const float4* pos, int n);   // the forward a training step starts with, for callers that merge it with another batch                                                                  // device scalar the step's loss is added to (may be null)
int dqn_train_prepare(DqnDev& d, DqnTrain& t, int n, cudaStream_t s);                     // allocations, first transposes, side stream: call before capturing a step
int dqn_alloc(DqnDev& d, int k_in);
void dqn_free(DqnDev& d);
int dqn_upload(DqnDev& d, const DqnHost& h, const float* vertices, cudaStream_t s);       // copies parameters, derives operands
int dqn_download(const DqnDev& d, DqnHost& h, cudaStream_t s);
int dqn_refresh_operands(DqnDev& d, cudaStream_t s);                                       // after the parameters changed on the device
int dqn_forward(const DqnDev& d, const DqnFwdParams& p, cudaStream_t s);                   // launches k_dqn_forward
int dqn_set_smem_limit();
void dqn_init_glorot(DqnHost& h, int k_in, uint32_t seed);
int dqn_load_text(DqnHost& h, const char* path, std::string& err);
int dqn_save_text(const DqnHost& h, const char* path, std::string& err);

}  // namespace rlpt
