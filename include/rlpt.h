/* include/rlpt.h -- C ABI of the B200-native reinforcement-learned path tracing hot path.
 *
 * This is the drop-in boundary for the hot path of callumPearce/Reinforcement-Light-Rays-Pathtracer
 * (GPU_Rendering_Engine/Source, shorthand G/). The reference has no FFI of its own: its "API" is the set of host
 * C++ types and __global__ symbols G/main.cu drives. Each entry point below names the reference interface it
 * replaces (file:line). The C++ mirror of the reference's host types (Scene, Camera, RadianceMap, SDLScreen) lives
 * in reinforcement-light-rays-pathtracer_b200/host/ and is written purely against this header; INTEGRATION.md
 * shows the binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero rlpt_status
 * otherwise, with the text available from rlpt_last_error(); nothing here ever calls exit() (the reference's
 * checkCudaErrors does: G/utils/cuda_helpers.cu:6-14). One context drives one GPU on its own CUDA stream; calls on
 * one context must be serialised by the caller. Host pointers unless a parameter is named d_*.
 * There is no CPU fallback: every call that computes needs the GPU and fails loudly without one.
 *
 * Image layout: pixel index = x*height + y (x-major), as in the reference (G/path_tracing/default_path_tracing.cu:13).
 */
#ifndef RLPT_H
#define RLPT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLPT_GRID_RESOLUTION 12          /* G/constants/radiance_volumes_settings.h:9 (compile-time in both engines) */
#define RLPT_GRID_CELLS 144

typedef struct rlpt_ctx rlpt_ctx;

enum rlpt_status {
    RLPT_OK = 0,
    RLPT_ERR_CUDA = 1,        /* a CUDA runtime call failed; rlpt_last_error() carries "CUDA error = <n> at file:line 'expr'" */
    RLPT_ERR_ARG = 2,         /* bad argument / call order (e.g. render before scene upload) */
    RLPT_ERR_IO = 3,
    RLPT_ERR_COLLECTIVE = 4,  /* the all-reduce hook reported failure */
    RLPT_ERR_UNSUPPORTED = 5
};

enum rlpt_traversal { RLPT_TRAVERSAL_AUTO = 0, RLPT_TRAVERSAL_BVH = 1, RLPT_TRAVERSAL_BRUTE = 2 };
enum rlpt_hit_type { RLPT_HIT_NOTHING = 0, RLPT_HIT_AREA_LIGHT = 1, RLPT_HIT_SURFACE = 2 };   /* G/rays/ray.cuh:30-34 */

/* Run-time form of the reference's compile-time settings (G/constants/ headers). rlpt_config_default() fills in the
 * reference's committed values, with the BASELINE.json resolution. */
typedef struct rlpt_config {
    int32_t width;               /* SCREEN_WIDTH            image_settings.h:9   */
    int32_t height;              /* SCREEN_HEIGHT = FOCAL_LENGTH  image_settings.h:10-11 */
    int32_t spp;                 /* SAMPLES_PER_PIXEL per frame   monte_carlo_settings.h:9 */
    int32_t max_bounces;         /* MAX_RAY_BOUNCES         monte_carlo_settings.h:8  */
    float env_light;             /* ENVIRONMENT_LIGHT       monte_carlo_settings.h:10 */
    float area_per_sample;       /* AREA_PER_SAMPLE         radiance_volumes_settings.h:12 */
    float max_dist;              /* MAX_DIST                radiance_volumes_settings.h:15 */
    float initial_radiance;      /* INITIAL_RADIANCE        radiance_volumes_settings.h:16 */
    float radiance_threshold;    /* RADIANCE_THRESHOLD      radiance_volumes_settings.h:17 */
    uint32_t seed;               /* curand_init(1984, ...)  utils/cuda_helpers.cu:24 */
    int32_t traversal;           /* enum rlpt_traversal */
    int32_t rank;                /* sample partition: this context traces samples rank*spp .. rank*spp+spp-1 of every */
    int32_t world_size;          /*   global frame of spp*world_size samples (counter-based RNG, so the union over ranks is
                                      independent of the GPU count) */
} rlpt_config;

/* All-reduce hook: sum `count` elements of `d_buf` (device memory of this context's GPU) in place across all
 * ranks, ordered on `cuda_stream`. dtype: 0 = float32, 1 = uint32. Return 0 on success.
 * The library calls it once per training iteration on the Q-table accumulators (sum of targets, visit counts)
 * and once per rlpt_frame_allreduce on the frame buffer. Hosts wire it to NCCL (see host/nccl_hook.cpp, or
 * torch.distributed in rlpt/__init__.py). */
typedef int (*rlpt_allreduce_fn)(void* d_buf, uint64_t count, int dtype, void* cuda_stream, void* user);

const char* rlpt_last_error(void);
int rlpt_version(void);

/* --- context ------------------------------------------------------------------------------------------- */
/* replaces: the implicit device 0 / default stream of G/main.cu (no cudaSetDevice anywhere in the reference) */
int rlpt_ctx_create(int device, rlpt_ctx** out);
int rlpt_ctx_destroy(rlpt_ctx* ctx);
int rlpt_sync(rlpt_ctx* ctx);
int rlpt_stream(rlpt_ctx* ctx, void** cuda_stream);
int rlpt_config_default(rlpt_config* cfg);
int rlpt_config_set(rlpt_ctx* ctx, const rlpt_config* cfg);      /* replaces G/constants/ headers #defines */
int rlpt_config_get(rlpt_ctx* ctx, rlpt_config* cfg);
int rlpt_set_allreduce(rlpt_ctx* ctx, rlpt_allreduce_fn fn, void* user);

/* --- scene --------------------------------------------------------------------------------------------- */
/* replaces: the cudaMalloc/cudaMemcpy upload of Scene, Surface[] and AreaLight[] with pointer patching
 * (G/main.cu:161-186). Triangles are 9 floats (v0,v1,v2) as the reference's Surface/AreaLight hold them
 * (G/objects/triangle.cuh:19-22); colours are Material::diffuse_c / AreaLight::diffuse_p. Normals
 * (G/objects/triangle.cu:67-76) and luminances (G/objects/material.cu:4-14) are derived here.
 * Emits SoA float4 triangle buffers and builds the BVH on the GPU. */
int rlpt_scene_upload(rlpt_ctx* ctx, const float* surface_v, const float* surface_rgb, int n_surfaces,
                      const float* light_v, const float* light_rgb, int n_lights);
int rlpt_scene_info(rlpt_ctx* ctx, int* n_surfaces, int* n_lights, int* bvh_nodes, int* bvh_depth);
/* device-built BVH, for inspection/tests: nodes as 16 floats each (see DESIGN.md "BVH node") */
int rlpt_scene_bvh_download(rlpt_ctx* ctx, float* nodes16, int max_nodes);
/* the 4-wide tree the tracing kernels walk (the binary tree above collapsed on the GPU; new work, the reference scans every
 * primitive, G/rays/ray.cu:16-36): 28 floats per node = lo.x, hi.x, lo.y, hi.y, lo.z, hi.z of the four children + four links
 * (>= 0 node, < 0 ~(first record << 3 | records), -1 = no records = unused slot); record_gid[i] = primitive id of leaf-order record i */
int rlpt_scene_bvh4_info(rlpt_ctx* ctx, int* nodes, int* depth, int* leaf_max);
int rlpt_scene_bvh4_download(rlpt_ctx* ctx, float* nodes28, int max_nodes, int* record_gid, int max_records);

/* replaces: cudaMemcpy(device_camera, &camera, ...) each frame (G/main.cu:210,307); Camera{position,yaw_y,yaw_x} (G/camera.cuh:19-22) */
int rlpt_camera_set(rlpt_ctx* ctx, const float position[4], float yaw_y, float yaw_x);

/* --- closest hit (parity entry point) -------------------------------------------------------------------- */
/* replaces: Ray::Ray + Ray::closest_intersection (G/rays/ray.cu:6-36) for a batch of rays. `dir` is normalised
 * the way Ray::Ray does. Outputs per ray: type (enum rlpt_hit_type), index into surfaces or lights (-1 for a
 * miss) and t in the reference's units (direction scaled by SCREEN_HEIGHT, G/rays/ray.cu:53; 999999 for a miss).
 * `traversal` overrides the configured mode when non-zero. counters (may be NULL): [0] triangle tests, [1] box tests. */
int rlpt_closest_hit(rlpt_ctx* ctx, const float* org, const float* dir, int n_rays, int traversal,
                     int* type, int* index, float* t, unsigned long long* counters);
/* same on device buffers, asynchronous on the context stream (used by bench.py to time traversal alone) */
int rlpt_closest_hit_device(rlpt_ctx* ctx, const float* d_org, const float* d_dir, int n_rays, int traversal,
                            int* d_type, int* d_index, float* d_t, unsigned long long* d_counters);

/* --- radiance map (Expected-SARSA Q-table) --------------------------------------------------------------- */
/* replaces: RadianceMap::RadianceMap (G/radiance_volumes/radiance_map.cu:8-54: volume count per surface :60-67,
 * rand() sampling :72-84, kd-tree RadianceTree G/radiance_volumes/radiance_tree.cu:12-62,135-196) and its upload
 * (G/main.cu:274-289). Volumes and tree are identical to the reference's (same rand() stream, same std::sort);
 * device storage is SoA (DESIGN.md "Q-table layout"). */
int rlpt_radiance_map_build(rlpt_ctx* ctx);
/* wall-clock seconds of the last build: [0] volume sampling + kd-tree (host; the reference's RadianceMap constructor, G/radiance_volumes/radiance_map.cu:8-54),
 * [1] nearest-volume candidate cells (host), [2] uploads and the first CDF build (device), [3] total */
int rlpt_radiance_map_build_seconds(rlpt_ctx* ctx, double* seconds4);
int rlpt_radiance_map_info(rlpt_ctx* ctx, int* n_volumes, int* n_tree_nodes);
/* Peer-memory exchange for N > 1 ranks on one node (one process per GPU). Without it the Q accumulators are all-reduced
 * through the rlpt_set_allreduce hook and merged afterwards; with it ONE kernel per rank reduces its slice of the volumes
 * straight out of every rank's accumulators (P2P loads over NVLink), merges, rebuilds the CDFs and stores the results into
 * every rank's tables (P2P stores). Call on every rank after rlpt_radiance_map_build: export this rank's blob
 * (rlpt_p2p_blob_bytes() bytes of CUDA IPC handles), gather all ranks' blobs in rank order by any means, import them.
 * rlpt_config.rank / world_size select the slice. The reference has no multi-GPU code. */
int rlpt_p2p_blob_bytes(void);
int rlpt_p2p_export(rlpt_ctx* ctx, void* blob);
int rlpt_p2p_import(rlpt_ctx* ctx, const void* blobs, int world_size);
/* back to the all-reduce hook (every rank must switch at the same frame boundary) */
int rlpt_p2p_close(rlpt_ctx* ctx);
/* flattened kd-tree exactly as the reference's std::vector<RadianceTreeElement> (G/radiance_volumes/radiance_tree.cuh:19-27) */
int rlpt_radiance_map_tree(rlpt_ctx* ctx, int* dim, int* leaf, unsigned* left, unsigned* right, float* data, float* pos3, float* nrm3);
/* replaces: RadianceMap::find_closest_radiance_volume_iterative (radiance_map.cu:150-203) for a batch of points */
int rlpt_radiance_map_find_closest(rlpt_ctx* ctx, const float* pos, const float* nrm, int n, int* volume_index);
/* replaces: writing RadianceVolume::radiance_grid on the host before upload (tests, checkpoints) */
int rlpt_radiance_map_set_q(rlpt_ctx* ctx, const float* q, const unsigned* visits /* may be NULL */);
/* replaces: update_radiance_volume_distributions<<<>>> (G/path_tracing/reinforcement_path_tracing.cu:6-13) */
int rlpt_radiance_map_update_distributions(rlpt_ctx* ctx);
/* replaces: cudaMemcpy of RadianceVolume[] back to the host (G/main.cu:372,389). Any pointer may be NULL.
 * q, cdf: n_volumes*144 floats; visits: n_volumes*144; irradiance: n_volumes; pos3/nrm3: 3*n_volumes; surface: n_volumes */
int rlpt_radiance_map_download(rlpt_ctx* ctx, float* q, float* cdf, unsigned* visits, float* irradiance,
                               float* pos3, float* nrm3, int* surface);
/* per-iteration accumulators (sum of TD targets, visit counts) before the merge; n_volumes*144 each */
int rlpt_radiance_map_delta_download(rlpt_ctx* ctx, float* target_sum, unsigned* count);
/* replaces: RadianceMap::save_q_vals_to_file (radiance_map.cu:237-268): "144\n" then "px py pz q0 .. q143" per volume */
int rlpt_radiance_map_save_q(rlpt_ctx* ctx, const char* path);
/* loader back into the renderer (the reference has none: SURVEY section 8f.2) */
int rlpt_radiance_map_load_q(rlpt_ctx* ctx, const char* path);

/* --- rendering ------------------------------------------------------------------------------------------- */
/* replaces: the method-0 frame loop, draw_default_path_tracing<<<>>> (G/main.cu:207-244,
 * G/path_tracing/default_path_tracing.cu:7-88). Each frame adds cfg.spp samples per pixel to the frame buffer. */
int rlpt_render_default(rlpt_ctx* ctx, int frames);
/* replaces: the method-1 frame loop, draw_reinforcement_path_tracing<<<>>> + update_radiance_volume_distributions<<<>>>
 * (G/main.cu:301-364, G/path_tracing/reinforcement_path_tracing.cu:15-120). One frame = one training iteration:
 * trace cfg.spp samples per pixel (TD targets accumulated with warp-aggregated atomics), all-reduce the
 * accumulators if a hook is set, merge into Q, rebuild CDFs. */
int rlpt_render_sarsa(rlpt_ctx* ctx, int frames);
/* the two halves of a SARSA frame, for hosts that want to interleave their own work */
int rlpt_sarsa_trace(rlpt_ctx* ctx);
int rlpt_sarsa_merge(rlpt_ctx* ctx);
/* render with the learned distributions frozen (no TD accumulation, no merge): train once, render many */
int rlpt_render_sarsa_frozen(rlpt_ctx* ctx, int frames);
/* replaces: the greedy debug samplers the reference swaps in by hand -- RadianceVolume::sample_max_direction_from_radiance_distribution
 * (G/radiance_volumes/radiance_volume.cu:248-278) for the radiance-volume tracer, sample_max_direction (G/deep_learning/nn_rendering_helpers.cu:492-553)
 * for the pretrained Neural-Q tracer: the next direction is a random point of the cell with the largest Q. on != 0 switches rlpt_render_sarsa,
 * rlpt_render_sarsa_frozen and rlpt_render_pretrained to it until switched off. */
int rlpt_set_max_direction(rlpt_ctx* ctx, int on);
/* replaces: the VORONOI debug view (G/main.cu:413-470, G/path_tracing/voronoi_trace.cu:4-45): one camera sample per pixel,
 * surface hits painted with the colour of their nearest radiance volume, everything else white. Sets the frame buffer. */
int rlpt_render_voronoi(rlpt_ctx* ctx);

/* replaces: cudaMemset(device_buffer, 0, ...) (G/main.cu:241,359) -- and resets the sample counter */
int rlpt_frame_reset(rlpt_ctx* ctx);
/* sums frame buffers and sample counts across ranks through the all-reduce hook (no-op without one) */
int rlpt_frame_allreduce(rlpt_ctx* ctx);
/* replaces: cudaMemcpy(host_buffer, device_buffer, ...) (G/main.cu:232,349): mean radiance, 3*width*height floats */
int rlpt_frame_download(rlpt_ctx* ctx, float* rgb);
/* replaces: the PutPixelSDL loop (G/main.cu:235-239, G/sdl/sdl_screen.cpp:96-108): ARGB8888, buffer index y*width+x */
int rlpt_frame_download_argb(rlpt_ctx* ctx, uint32_t* argb);
/* replaces: SDLScreen::SDL_SaveImage (G/sdl/sdl_screen.cpp:60-66): 32-bpp BITMAPV4HEADER BMP */
int rlpt_frame_save_bmp(rlpt_ctx* ctx, const char* path);

/* replaces: the path-length / zero-contribution readbacks and printfs (G/main.cu:223-229,322-339).
 * Totals since the last rlpt_stats_reset: paths traced, sum of path lengths, zero-contribution paths,
 * ray casts, and device seconds spent inside render calls (CUDA events on the context stream). */
typedef struct rlpt_stats_t {
    double paths;
    double path_length_sum;
    double zero_contribution_paths;
    double ray_casts;
    double device_seconds;
    double frames;
    double kernel_launches;
    double triangle_tests;        /* ray-triangle solves executed by the tracing kernels (72 flop each) */
    double box_tests;             /* ray-AABB slab tests executed (18 flop each) */
    double trace_seconds;         /* device seconds inside the per-bounce tracing kernels (CUDA events) */
    double merge_seconds;         /* device seconds inside all-reduce + Q merge + CDF rebuild */
    double kd_fallbacks;          /* nearest-volume queries the candidate cells could not decide (answered by the reference's kd search) */
    double isect_seconds;         /* device seconds inside k_isect launches (event pairs on the launching streams, taken on every fourth frame and scaled to all launches) and their number */
    double isect_launches;
    double shade_seconds;         /* the same for k_shade */
    double shade_launches;
    double tail_seconds;          /* the same for the run-to-completion k_bounce launches */
    double tail_launches;
    double dqn_forward_seconds;   /* Neural-Q tracers: device seconds inside the per-bounce k_dqn_forward launches over all live rays (event pairs), */
    double dqn_forward_launches;  /*   their number, */
    double dqn_forward_rays;      /*   and the rays they evaluated (dense-equivalent flop per ray: SURVEY 8d) */
    double train_steps;           /* optimiser steps taken by rlpt_render_neuralq (one per batch and bounce) */
    double train_seconds;         /* device seconds inside them (next-state forward, TD targets, forward + backward, Adam) */
} rlpt_stats_t;
int rlpt_stats(rlpt_ctx* ctx, rlpt_stats_t* out);
int rlpt_stats_reset(rlpt_ctx* ctx);

/* --- Neural-Q network (DQN) ------------------------------------------------------------------------------ */
/* The network of N/dq_network.cu:8-49: input = every scene vertex minus the query point (9 floats per triangle,
 * G/deep_learning/nn_rendering_helpers.cu:280-298), 200-300-200-144 with ReLU after every layer, evaluated on the
 * tensor cores (bf16 operands, fp32 accumulation; layer 1 in fp32). Parameters are the 8 DyNet blocks in file order
 * (W1 b1 W2 b2 W3 b3 W4 b4), each W row-major [out][in]; total 200*k_in + 200 + 60000 + 300 + 60000 + 200 + 28800 + 144 floats. */
/* replaces: Scene::vertices as the network's constant input (G/main.cu:171-176). Default when not called: the uploaded
 * triangles, surfaces then lights, in v0 v1 v2 order. Call after rlpt_scene_upload. */
int rlpt_dqn_set_vertices(rlpt_ctx* ctx, const float* vertices, int count);
/* replaces: DQNetwork::initialize + DyNet's default Glorot initialiser (G/deep_learning/neural_q_pathtracer.cu:44-50) */
int rlpt_dqn_init(rlpt_ctx* ctx, uint32_t seed);
/* replaces: dynet::TextFileLoader::populate / TextFileSaver::save (neural_q_pathtracer.cu:55-59,191-196); same text format */
int rlpt_dqn_load_text(rlpt_ctx* ctx, const char* path);
int rlpt_dqn_save_text(rlpt_ctx* ctx, const char* path);
int rlpt_dqn_param_count(rlpt_ctx* ctx, int* count, int* k_in);
int rlpt_dqn_set_params(rlpt_ctx* ctx, const float* params, int count);
int rlpt_dqn_get_params(rlpt_ctx* ctx, float* params, int count);
/* replaces: convert_vertices_to_point_coord_system + network_inference + forward for a batch of query points
 * (neural_q_pathtracer.cu:292-325). q: n*144 floats, row-major [point][action]. */
int rlpt_dqn_forward(rlpt_ctx* ctx, const float* pos3, int n, float* q);

/* replaces: one optimiser step of the reference's training loop (G/deep_learning/neural_q_pathtracer.cu:476-512): forward
 * of the states pos3 with training=true, loss = sum_b (target_b - Q(s_b)[action_b])^2 (dynet::pick, pow, sum_batches),
 * backward, AdamTrainer::update (DyNet defaults). apply_update = 0 computes loss and gradients only. loss may be NULL. */
int rlpt_dqn_train_batch(rlpt_ctx* ctx, const float* pos3, const uint32_t* actions, const float* targets, int n, int apply_update, float* loss);
/* gradients of the last rlpt_dqn_train_batch in parameter order (tests) */
/* replaces: the training step of the offline trainer (NN_Q_Value_Trainer/Source/main.cu:67-135): fit the network to saved Q
 * tables. targets144 = n*144 floats (ray-major); loss = sum over the batch of the squared distance over all 144 outputs; one
 * Adam step when apply_update != 0. */
int rlpt_dqn_train_supervised(rlpt_ctx* ctx, const float* pos3, const float* targets144, int n, int apply_update, float* loss);
int rlpt_dqn_get_grads(rlpt_ctx* ctx, float* grads, int count);

/* replaces: PretrainedPathtracer(frames, batch, screen, scene, camera, ...) (G/deep_learning/pre_trained_pathtracer.cu:10-491):
 * path tracing with directions importance-sampled from the network's Q values, no learning. Each frame adds cfg.spp
 * samples per pixel to the frame buffer. */
int rlpt_render_pretrained(rlpt_ctx* ctx, int frames);

/* replaces: NeuralQPathtracer(frames, batch_size, screen, scene, camera, argc, argv) (G/deep_learning/neural_q_pathtracer.cu:5-600):
 * train the network while rendering. Each frame makes cfg.spp passes over all pixels; every bounce of a pass performs
 * ceil(width*height / batch) sequential optimiser steps (batch = 4096 in G/main.cu:116-118). With an all-reduce hook and
 * world_size > 1 the gradients of every step are summed across ranks. */
int rlpt_render_neuralq(rlpt_ctx* ctx, int frames, int batch);
/* replaces: the compile-time training constants EPSILON_START / EPSILON_DECAY / EPSILON_MIN (G/constants/deep_learning_settings.h:5-7) and the
 * learning rate handed to DyNet's AdamTrainer (G/deep_learning/neural_q_pathtracer.cu:44-53, DyNet default 0.001). epsilon_start also resets the
 * running epsilon. learning_rate 0 freezes the network (the tracer still runs every optimiser step): what the same-path parity test uses. */
int rlpt_neuralq_set_hyper(rlpt_ctx* ctx, float learning_rate, float epsilon_start, float epsilon_decay, float epsilon_min);
/* the loss summed over the last frame ("loss" of nn_training_stats.txt, neural_q_pathtracer.cu:578-583) */
int rlpt_neuralq_last_loss(rlpt_ctx* ctx, double* loss);

/* --- measurement helpers (bench.py) ---------------------------------------------------------------------- */
/* FP32 FMA microbenchmark on the context's GPU: returns achieved TFLOP/s (2 flop per FMA). */
int rlpt_measure_fp32_peak(rlpt_ctx* ctx, double* tflops);
/* dump the rays the wavefront casts at bounce `bounce` of the next traced frame (device capture -> host), at most max_rays */
int rlpt_capture_rays(rlpt_ctx* ctx, int method, int bounce, float* org, float* dir, int max_rays, int* n_captured);

#ifdef __cplusplus
}
#endif
#endif /* RLPT_H */
