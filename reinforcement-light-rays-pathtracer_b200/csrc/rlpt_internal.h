// rlpt_internal.h -- device-side data layout shared by the kernels and the C-ABI host code (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rlpt_device.cuh"

namespace rlpt {

// ---- scene in HBM: SoA float4 buffers (the .obj loader / scene upload emits these; DESIGN.md "Scene layout")
// Primitive id ("gid"): surfaces 0..n_surf-1 then lights n_surf..n_tri-1, the order Ray::closest_intersection scans
// them in (G/rays/ray.cu:22-35), so "lowest gid wins a tie" is the reference's tie rule.
//   tri[3*gid+0] = (v0.x, v0.y, v0.z, e1.x)   tri[3*gid+1] = (e1.y, e1.z, e2.x, e2.y)   tri[3*gid+2] = (e2.z, T1, 0, 0)
//   shade[4*gid+0] = (N.x, N.y, N.z, luminance/pi)                [lights: luminance]
//   shade[4*gid+1] = (T.x, T.y, T.z, as_float(normal class))
//   shade[4*gid+2] = (B.x, B.y, B.z, 0)
//   shade[4*gid+3] = (diffuse_c / pi  rgb, 0)                     [lights: diffuse_p rgb]
//   bvh4[7*node+0..6]: the 4-wide traversal tree, see rlpt_bvh.cu: six plane records (lo.x, hi.x, lo.y, hi.y, lo.z, hi.z of the four
//                      children) + four child links; tri4 = the triangle records in LEAF order (a leaf = a run of 1..BVH4_LEAF_MAX
//                      records) with the primitive id in the third float4's z
struct SceneDev {
    const float4* tri;
    const float4* shade;
    const float4* bvh4;
    const float4* tri4;
    int n_tri, n_surf, n_light, n_nodes4;
    int brute;          // 1: scan all primitives in gid order from shared memory; 0: BVH traversal
    int staged;         // 1: the scene is copied into shared memory by every CTA (brute: tri, shade, scan units; BVH: tri4, shade, bvh4); 0: read-only path / L1
    // conservative pre-test of the brute-force scan (rlpt_device.cuh, unit_candidates): pairs of triangles that form a
    // parallelogram are tested together, 4 float4 per pair: (v0, e1.x) (e1.yz, e2.xy) (e2.z, n) (pu, pv, ps, -) with n = e1 x e2
    // and (pu, pv, ps) placing the second triangle in the first one's (u, v). slot_gid[slot] = primitive id: slots 2u, 2u+1 are
    // pair u, the unpaired primitives follow (n_tri slots in all, padded to a multiple of 4).
    const float4* scan; const int* slot_gid; int n_units; int n_items; int bundle;     // bundle: camera rays are pre-tested once per warp (closest_hit_bundle)   // n_items >= n_units records: the unpaired triangles follow the pairs (bundle pre-test)
    float k1, k2, k3, vmax;   // error-bound coefficients of the pre-test (host: choose_traversal), largest |vertex coordinate|
    int det_small;      // 1 when SCREEN_HEIGHT * max |e1| |e2| < 2^23: no determinant of this scene can reach the range where tri_candidate's sign test needs its guard
};

// ---- radiance map in HBM (SoA; the reference keeps one 1832-byte AoS record per volume, radiance_volume.cuh:40-49)
struct RadianceDev {
    const float4* kd_inner;   // (split, as_float(left), as_float(right), as_float(dim)); leaf children = KD_LEAF | volume
    const float4* vol_posn;   // (position, as_float(normal class))
    const int* vol_surface;
    float* q;                 // [n_vol][144]  Q(x, omega)                      radiance_grid
    float* cdf;               // [n_vol][144]  inclusive CDF frozen per frame  radiance_distribution
    float* cdf_rows;          // [n_vol][12]   cdf[12 j + 11]: first level of the two-level sector search
    uint32_t* visits;         // [n_vol][144]
    float* irradiance;        // [n_vol]       sum_k Q_k cos_k lum/pi           irradiance_accum
    float* acc_sum;           // [n_vol][144]  sum of TD targets this iteration
    uint32_t* acc_cnt;        // [n_vol][144]  visits this iteration
    int n_vol, n_inner;
    uint32_t root;            // KD child word of the root
    float root_px, root_py, root_pz;   // radiance_array[0].position as the reference initialises its search with
    float within_abs;         // exact float form of the reference's pow(delta,2) < MAX_DIST test (rlpt_device.cuh, kd_find)
    // candidate-cell front end of the nearest-volume search (rlpt_device.cuh, vcell_find)
    VCells vc;
    const int4* vc_table;     // [vc.mask + 1] (cell, class, first candidate group, groups of 4); cell = -1: empty
    const float4* vc_cand;    // (position, as_float(volume index)); lists padded to groups of 4 with volume -1
    const int4* vx_table;     // second level (rlpt_device.cuh, vext_find): same slot format, 4th word = candidates
    const float4* vx_cand;    // 3 float4 per candidate: (position, volume) (kd-cell lo, hi.x) (hi.yz, -, -)
    uint32_t vx_mask;
};

// ---- wavefront path state, SoA, one slot per live path (two queues, ping-pong per bounce)
//   o   = (origin, as_float(pixel))       d = (direction, BRDF luminance/pi of the surface the ray left)
//   thr = (throughput rgb, as_float(volume << 8 | sector))       meta = sample << 8 | bounce
struct PathQueue { float4* o; float4* d; float4* thr; uint32_t* meta; };

// per-frame, per-lane values; passed to the kernels by value
struct FrameDyn {
    uint32_t sample_base; int learn;
    float cam_x, cam_y, cam_z, cy, sy, cx, sx; int rotated;
    int capture_bounce, capture_max;
    int max_dir;                    // 1: greedy debug sampling (sample_max_direction_from_radiance_distribution / sample_max_direction)
};

// One lane = one slice of a frame's samples traced on its own stream with its own queues, so that the latency-bound
// tail bounces of one slice overlap the wide early bounces of the next (DESIGN.md "Wavefront").
struct FrameParams {
    SceneDev scene;
    RadianceDev rm;
    PathQueue q[2];
    float2* hit;                    // split pipeline: (t, as_float(primitive id)) per slot of the current queue
    int* counts;                    // live paths entering each bounce: sub-queue k of bounce b at [(b * NSUB + k) * COUNT_STRIDE] (the Neural-Q tracers keep one queue: [b])
    int sub_cap;                    // slots per sub-queue
    int* cursor;                    // same layout as counts: next unclaimed ray of each (bounce, sub-queue), for k_isect_bvh's dynamic ray fetch
    float4* accum;                  // [W*H] radiance sums (rgb) + sample count in w
    unsigned long long* stats;      // [0] path-length sum, [1] zero-contribution paths, [2] terminated paths, [3] tri tests, [4] box tests, [5] kd-search fallbacks
    float4* capture_o; float4* capture_d; int* capture_n;
    int width, height, spp, max_bounces;     // spp = samples per pixel traced by this lane
    uint32_t seed;
    float env;
};

constexpr int BLOCK = 256;
constexpr int NSUB = 32;            // sub-queues per lane (rlpt_kernels.cu, "Sub-queues"); the tracing grids are multiples of it
constexpr int COUNT_STRIDE = 8;     // ints between two sub-queue counters: one 32-byte sector each

// host-visible launchers (rlpt_kernels.cu)
void launch_primary(const FrameParams& p, const FrameDyn& dyn, int method, int grid, size_t smem, cudaStream_t s);
void launch_bounce(const FrameParams& p, const FrameDyn& dyn, int method, int bounce, int grid, size_t smem, cudaStream_t s);
void launch_tail(const FrameParams& p, const FrameDyn& dyn, int method, int bounce, int grid, size_t smem, cudaStream_t s);
void launch_isect(const FrameParams& p, const FrameDyn& dyn, int bounce, int grid, size_t smem, cudaStream_t s);
void launch_shade(const FrameParams& p, const FrameDyn& dyn, int method, int bounce, int grid, cudaStream_t s);
// Neural-Q training tracer state (NeuralQPathtracer, G/deep_learning/neural_q_pathtracer.cu:76-96): one slot per pixel,
// every ray takes part in every bounce (terminated rays are re-seeded on the geometry and keep generating training data)
struct NqTrainState {
    float4* loc;        // (state position, as_float(surface id or -1))      ray_locations + ray_normals
    float4* sloc;       // the state the action was taken in (network input of the training step)
    float4* dir;        // direction to trace                                 ray_directions
    float4* thr;        // throughput rgb                                     ray_throughputs
    uint32_t* state;    // 0 alive, 1 terminated this bounce, 2 learning only  ray_states
    float* reward; float* discount;                                        // ray_rewards, ray_discounts
    uint32_t* action;   // chosen grid cell                                   directions_host
    int* alive;         // [max_bounces + 2] alive (state 0) rays entering each bounce
    int n;              // width * height
};
void launch_nqt_init(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, cudaStream_t s);
void launch_nqt_sample(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, int bounce, const float* q, int q_stride, float epsilon, cudaStream_t s);
void launch_nqt_trace(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, int bounce, int grid, size_t smem, cudaStream_t s);
void launch_nqt_respawn(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, int bounce, cudaStream_t s);
void launch_add_scalar(float* dst, const float* src, cudaStream_t s);
// Neural-Q wavefront (rlpt_kernels.cu): trace without sampling (the direction comes from the network), and the sampler
void launch_nq_trace(const FrameParams& p, const FrameDyn& dyn, int bounce, int grid, size_t smem, cudaStream_t s);
void launch_nq_sample(const FrameParams& p, const FrameDyn& dyn, int bounce, const float* q, int q_stride, float epsilon, uint32_t* action_out, int grid, cudaStream_t s);
void launch_merge(const RadianceDev& rm, const float* surf_lum_over_pi, float threshold, int rebuild_only, cudaStream_t s);
// Peer-memory exchange (one process per GPU of one node, buffers opened through CUDA IPC): the same arrays on every rank.
// flags: [2][MAX_PEERS] unsigned per rank -- [0][r] = rank r's accumulators of epoch e are complete, [1][r] = rank r has
// stored its slice of epoch e into my tables (and is done with my accumulators).
constexpr int MAX_PEERS = 8;
struct PeerTables {
    float* acc_sum[MAX_PEERS]; uint32_t* acc_cnt[MAX_PEERS];
    float* q[MAX_PEERS]; float* cdf[MAX_PEERS]; float* cdf_rows[MAX_PEERS]; uint32_t* visits[MAX_PEERS]; float* irradiance[MAX_PEERS];
    unsigned* flags[MAX_PEERS];
    unsigned* error;            // this rank's error word (device memory): raised when a wait for a peer times out
    int world, rank;
};
void launch_merge_p2p(const RadianceDev& rm, const PeerTables& pt, const float* surf_lum_over_pi, float threshold, unsigned epoch, unsigned* done_counter, cudaStream_t s);
void launch_closest_hit(const SceneDev& sc, const float* org, const float* dir, int n, float H, int* type, int* index, float* t,
                        unsigned long long* counters, int* cursor, size_t smem, cudaStream_t s);
size_t bvh4_scratch_bytes();
void launch_find_closest(const RadianceDev& rm, const SceneDev& sc, const float* pos, const float* nrm, int n, int* out, cudaStream_t s);
void launch_voronoi(const FrameParams& p, const FrameDyn& dyn, int grid, size_t smem, cudaStream_t s);
void launch_frame_mean(const float4* accum, float* rgb, int n, cudaStream_t s);
void launch_pack_argb(const float4* accum, uint32_t* argb, int width, int height, cudaStream_t s);
void launch_fp32_peak(float* out, int iters, int grid, cudaStream_t s);
size_t scene_smem_bytes(const SceneDev& sc);
void upload_cell_cos(const float* cos144);
int kernels_set_smem_limit(size_t bytes);
void kernels_resident_ctas(size_t isect_smem, int brute, int staged, int* isect_per_sm, int* shade_per_sm);

// rlpt_bvh.cu: builds the BVH on the GPU from the tri buffer; returns node count and depth; d_bvh is allocated by the callee
// The binary tree (d_bvh, one primitive per leaf; kept for rlpt_scene_bvh_download) is collapsed on the GPU into the 4-wide tree the
// kernels walk (d_bvh4, d_tri4); leaf_max = primitives per leaf of the wide tree (1..BVH4_LEAF_MAX).
constexpr int BVH4_LEAF_MAX = 2;
constexpr int BVH4_EMPTY = -1;                      // child link of an unused slot: reads as a leaf of zero records, so even a NaN ray (every slab test passes) finds nothing in it
int bvh_build_gpu(const float4* d_tri, int n_tri, int leaf_max, float4** d_bvh, int* n_nodes, int* depth, float4** d_bvh4, int* n_nodes4, int* depth4, float4** d_tri4, cudaStream_t s);

}  // namespace rlpt
