// rlpt_capi.cu -- the extern "C" boundary declared in include/rlpt.h: context, scene upload (SoA + GPU BVH), radiance
// map lifecycle, the per-frame launch sequences, frame buffer and statistics. Host code only; every number the hot
// path produces comes from the kernels in rlpt_kernels.cu / rlpt_bvh.cu. No CPU fallback exists: without a usable
// GPU rlpt_ctx_create fails and nothing else can be called.
#include "../../include/rlpt.h"
#include "rlpt_internal.h"
#include "rlpt_radiance_host.h"
#include "rlpt_dqn.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <array>
#include <chrono>

using namespace rlpt;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

// same text as the reference's check_cuda (G/utils/cuda_helpers.cu:6-14), but returned instead of exit(99)
#define CK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { char b_[512]; \
    snprintf(b_, sizeof b_, "CUDA error = %u at %s:%d '%s' (%s)", (unsigned)e_, __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
    return fail(RLPT_ERR_CUDA, b_); } } while (0)

struct rlpt_ctx {
    int device = 0; int n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    rlpt_config cfg{};
    rlpt_allreduce_fn allreduce = nullptr; void* allreduce_user = nullptr;
    // scene
    bool have_scene = false;
    int n_surf = 0, n_light = 0, bvh_depth = 0;
    float4 *d_tri = nullptr, *d_shade = nullptr, *d_bvh = nullptr, *d_bvh4 = nullptr, *d_tri4 = nullptr, *d_scan = nullptr; int n_nodes = 0, bvh4_depth = 0, bvh_leaf_max = 2; int* d_chit_cursor = nullptr; int* d_scan_gid = nullptr; float* d_surf_lum_over_pi = nullptr;
    std::vector<float> h_surf_v, h_surf_nrm, h_surf_rgb, h_light_v, h_light_rgb;
    std::vector<int> h_surf_class;
    SceneDev scene{};
    // camera / per-frame
    float cam[3] = { 0.f, 0.f, -3.f }; float yaw_y = 0.f, yaw_x = 0.f; int max_dir = 0;
    // radiance map
    bool have_rmap = false; double rmap_build_s[4] = { 0.0, 0.0, 0.0, 0.0 };     // volumes + kd-tree (host), candidate cells (host), uploads + first CDF build (device), total
    std::vector<HostVolume> h_vol; std::vector<HostTreeElement> h_tree;
    float4 *d_kd = nullptr, *d_posn = nullptr; int* d_vol_surface = nullptr;
    int4* d_vc_table = nullptr; float4* d_vc_cand = nullptr; int4* d_vx_table = nullptr; float4* d_vx_cand = nullptr; float vx_built_within = 0.f; float vc_built_accept = 0.f; size_t vc_keys = 0, vc_listed = 0;   // nearest-volume candidate cells
    float *d_q = nullptr, *d_cdf = nullptr, *d_cdf_rows = nullptr, *d_irr = nullptr, *d_acc_sum = nullptr; uint32_t *d_visits = nullptr, *d_acc_cnt = nullptr;
    RadianceDev rm{};
    // peer-memory exchange (rlpt_p2p_export / rlpt_p2p_import): every rank's exchange buffers opened through CUDA IPC
    PeerTables peers{}; bool p2p_ready = false; unsigned p2p_epoch = 0; unsigned* d_p2p_flags = nullptr; unsigned* d_p2p_done = nullptr; unsigned* d_p2p_error = nullptr; bool p2p_used = false;
    std::vector<void*> p2p_opened;
    // Neural-Q network
    DqnHost dq_host; DqnDev dq; std::vector<float> dq_vertices; bool dq_vertices_custom = false;
    DqnTrain dq_train;
    NqTrainState nqt{}; int nqt_n = 0; float *d_nqt_qcur = nullptr, *d_nqt_qnext = nullptr, *d_nqt_targets = nullptr, *d_nqt_loss = nullptr; int nqt_batch = 0;
    // CUDA graph of one full-batch optimiser step + fixed staging buffers for the batch's slice of the ray arrays
    int nq_graphs = 1; cudaGraphExec_t nq_graph_exec = nullptr; int nq_graph_batch = 0;
    float nq_epsilon = 0.05f, nq_eps_decay = 0.01f, nq_eps_min = 0.05f, nq_lr = 1e-3f; double nq_loss_total = 0.0;     // EPSILON_START / DECAY / MIN (G/constants/deep_learning_settings.h:5-7)
    float* d_nq_q = nullptr; size_t nq_q_capacity = 0;           // Q values of the live paths, [144][capacity]
    // wavefront state
    struct Lane {
        cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; PathQueue q[2]{}; int* d_counts = nullptr; int* d_cursor = nullptr; float2* d_hit = nullptr;
        // the primary closest-hit pass needs nothing of the previous frame's results, only the hit buffer: it runs on `pre`
        // as soon as the previous frame's last k_shade has read that buffer, underneath that frame's tail kernel and merge
        cudaStream_t pre = nullptr; cudaEvent_t hit_free = nullptr, pre_done = nullptr; bool hit_free_valid = false;
        // live-path counts of recent frames, copied back asynchronously: the launch plan of a frame (where the per-bounce
        // launches stop and the run-to-completion kernel takes over) is read off the newest snapshot that has arrived
        static constexpr int SNAPS = 4;
        int* h_counts = nullptr; cudaEvent_t snap_ev[SNAPS] = {}; int snap_tag[SNAPS] = {}; uint64_t snap_seq = 0;   // tag = max_bounces the snapshot was taken under (0: none)
    };
    std::vector<Lane> lanes; size_t lane_capacity = 0; int counts_len = 0; int lane_spp = 0; int sub_cap = 0; cudaEvent_t ev_fork = nullptr;
    float4* d_accum = nullptr; int accum_pixels = 0;
    unsigned long long* d_stats = nullptr;
    float4 *d_cap_o = nullptr, *d_cap_d = nullptr; int* d_cap_n = nullptr; int cap_max = 0, cap_bounce = -1;
    size_t smem_bytes = 0; int grid = 148;
    int pipe_split = 1, pipe_tail = 32768, pipe_pre = 1; int per_sm_isect = 0, per_sm_shade = 0; size_t resident_smem = (size_t)-1;   // tracing pipeline (DESIGN.md "Wavefront"): split launches, run-to-completion threshold
    void* d_stage = nullptr; size_t stage_bytes = 0;     // device staging for frame downloads (kept across calls)
    uint64_t frames_done = 0;            // global frame counter: sample_base = (frames_done*world + rank)*spp
    double device_seconds = 0.0, frames_rendered = 0.0, launches = 0.0, trace_seconds = 0.0, merge_seconds = 0.0;
    std::vector<cudaEvent_t> phase_ev; size_t phase_used = 0;      // per-frame phase marks of the current render call
    // per-kernel device times (roofline denominators): event pairs around the k_isect / k_shade / tail launches, on the
    // streams they are launched on; resolved into the sums below once the work has finished (rlpt_stats / end of a render call)
    struct KPair { cudaEvent_t a, b; int kind; };
    std::vector<cudaEvent_t> kev_pool; size_t kev_used = 0; std::vector<KPair> kev_pending;
    double k_seconds[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 }, k_launches[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 }, k_all[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 }; double dqn_rays = 0.0;     // 0 k_isect, 1 k_shade, 2 run-to-completion k_bounce, 3 per-bounce k_dqn_forward, 4 optimiser steps of one bounce; timed launches / all launches
};

static void free_scene(rlpt_ctx* c) {
    cudaFree(c->d_tri); cudaFree(c->d_shade); cudaFree(c->d_bvh); cudaFree(c->d_bvh4); cudaFree(c->d_tri4); c->d_bvh4 = c->d_tri4 = nullptr; cudaFree(c->d_surf_lum_over_pi); cudaFree(c->d_scan); cudaFree(c->d_scan_gid); c->d_scan = nullptr; c->d_scan_gid = nullptr;
    c->d_tri = c->d_shade = c->d_bvh = nullptr; c->d_surf_lum_over_pi = nullptr; c->have_scene = false;
}
// Closes this rank's view of the peers. The flag arrays, the error word and the epoch counter live as long as the context: peers
// may still hold the flags mapped (a cudaFree under them would leave dangling IPC mappings), and the epoch stays monotonic so that
// a later export / import round (whose blobs carry every rank's epoch) resynchronises instead of restarting from 0.
static void p2p_close(rlpt_ctx* c) {
    for (void* p : c->p2p_opened) cudaIpcCloseMemHandle(p);
    c->p2p_opened.clear(); c->p2p_ready = false; c->peers = PeerTables{};
}
static void p2p_destroy(rlpt_ctx* c) {
    p2p_close(c);
    cudaFree(c->d_p2p_flags); cudaFree(c->d_p2p_done); cudaFree(c->d_p2p_error); c->d_p2p_flags = c->d_p2p_done = c->d_p2p_error = nullptr;
}
// the error word of the peer-memory merge (rlpt_kernels.cu, wait_flag): call after the stream has drained
static int p2p_check(rlpt_ctx* c) {
    if (!c->p2p_used || !c->d_p2p_error) return RLPT_OK;
    unsigned e = 0;
    if (cudaMemcpy(&e, c->d_p2p_error, sizeof e, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(RLPT_ERR_CUDA, "peer-memory exchange: cannot read the error word");
    c->p2p_used = false;
    if (!e) return RLPT_OK;
    cudaMemset(c->d_p2p_error, 0, sizeof(unsigned)); cudaMemset(c->d_p2p_done, 0, sizeof(unsigned));
    c->p2p_ready = false;                                              // broken until the ranks run rlpt_p2p_export / rlpt_p2p_import again
    return fail(RLPT_ERR_COLLECTIVE, "peer-memory exchange timed out waiting for a rank (asymmetric rlpt_sarsa_merge calls, a failed or restarted peer); "
                                     "radiance tables may be incomplete -- rebuild the exchange with rlpt_p2p_export / rlpt_p2p_import");
}
static void free_rmap(rlpt_ctx* c) {
    p2p_close(c);
    cudaFree(c->d_kd); cudaFree(c->d_posn); cudaFree(c->d_vol_surface); cudaFree(c->d_q); cudaFree(c->d_cdf); cudaFree(c->d_irr);
    cudaFree(c->d_acc_sum); cudaFree(c->d_visits); cudaFree(c->d_acc_cnt); cudaFree(c->d_cdf_rows); c->d_cdf_rows = nullptr;
    cudaFree(c->d_vc_table); cudaFree(c->d_vc_cand); c->d_vc_table = nullptr; c->d_vc_cand = nullptr;
    cudaFree(c->d_vx_table); cudaFree(c->d_vx_cand); c->d_vx_table = nullptr; c->d_vx_cand = nullptr;
    c->d_kd = c->d_posn = nullptr; c->d_vol_surface = nullptr; c->d_q = c->d_cdf = c->d_irr = c->d_acc_sum = nullptr; c->d_visits = c->d_acc_cnt = nullptr;
    c->have_rmap = false; c->rm = RadianceDev{};
}
static void free_lanes(rlpt_ctx* c) {
    for (auto& l : c->lanes) {
        for (int k = 0; k < 2; ++k) { cudaFree(l.q[k].o); cudaFree(l.q[k].d); cudaFree(l.q[k].thr); cudaFree(l.q[k].meta); }
        cudaFree(l.d_counts); cudaFree(l.d_cursor); cudaFree(l.d_hit); if (l.h_counts) cudaFreeHost(l.h_counts);
        for (auto& e : l.snap_ev) if (e) cudaEventDestroy(e);
        if (l.done) cudaEventDestroy(l.done); if (l.stream) cudaStreamDestroy(l.stream);
        if (l.hit_free) cudaEventDestroy(l.hit_free); if (l.pre_done) cudaEventDestroy(l.pre_done); if (l.pre) cudaStreamDestroy(l.pre);
    }
    c->lanes.clear(); c->lane_capacity = 0; c->counts_len = 0; c->lane_spp = 0;
}
static void nq_graph_reset(rlpt_ctx* c) { if (c->nq_graph_exec) { cudaGraphExecDestroy(c->nq_graph_exec); c->nq_graph_exec = nullptr; } c->nq_graph_batch = 0; }
static void free_nqt(rlpt_ctx* c) {
    nq_graph_reset(c);
    cudaFree(c->nqt.loc); cudaFree(c->nqt.sloc); cudaFree(c->nqt.dir); cudaFree(c->nqt.thr); cudaFree(c->nqt.state); cudaFree(c->nqt.reward); cudaFree(c->nqt.discount);
    cudaFree(c->nqt.action); cudaFree(c->nqt.alive); cudaFree(c->d_nqt_qcur); cudaFree(c->d_nqt_qnext); cudaFree(c->d_nqt_targets); cudaFree(c->d_nqt_loss);
    c->nqt = NqTrainState{}; c->nqt_n = 0; c->nqt_batch = 0; c->d_nqt_qcur = c->d_nqt_qnext = c->d_nqt_targets = c->d_nqt_loss = nullptr;
}
static void free_frame(rlpt_ctx* c) {
    free_lanes(c); free_nqt(c);
    cudaFree(c->d_accum); c->d_accum = nullptr; c->accum_pixels = 0;
}

// largest float f with (double)f*f < (double)max_dist: fabsf(delta) <= f is the reference's pow(delta,2) < max_dist
static float within_abs_of(float max_dist) {
    const double md = (double)max_dist;
    if (!(md > 0.0)) return -1.f;
    float f = (float)std::sqrt(md);
    while (f > 0.f && (double)f * (double)f >= md) f = std::nextafterf(f, 0.f);
    for (float n = std::nextafterf(f, INFINITY); (double)n * (double)n < md; n = std::nextafterf(f, INFINITY)) f = n;
    return f;
}

static int ensure_stage(rlpt_ctx* c, size_t bytes) {
    if (bytes <= c->stage_bytes) return RLPT_OK;
    cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0;
    CK(cudaMalloc(&c->d_stage, bytes)); c->stage_bytes = bytes;
    return RLPT_OK;
}
static float luminance3(const float* c) { float mx = std::max(c[2], std::max(c[0], c[1])), mn = std::min(c[2], std::min(c[0], c[1])); return 0.5f * (mx + mn); }

extern "C" {

const char* rlpt_last_error(void) { return g_err.c_str(); }
int rlpt_version(void) { return 100; }

int rlpt_config_default(rlpt_config* cfg) {
    if (!cfg) return fail(RLPT_ERR_ARG, "rlpt_config_default: null");
    cfg->width = 512; cfg->height = 512; cfg->spp = 32; cfg->max_bounces = 80; cfg->env_light = 0.f;
    cfg->area_per_sample = 0.001f; cfg->max_dist = 0.003f;
    cfg->initial_radiance = (1.f / (12.f * 12.f)) * 100.f; cfg->radiance_threshold = (1.f / (12.f * 12.f)) * 0.8f;
    cfg->seed = 1984u; cfg->traversal = RLPT_TRAVERSAL_AUTO; cfg->rank = 0; cfg->world_size = 1;
    return RLPT_OK;
}

int rlpt_ctx_create(int device, rlpt_ctx** out) {
    if (!out) return fail(RLPT_ERR_ARG, "rlpt_ctx_create: null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(RLPT_ERR_CUDA, std::string("rlpt_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (device < 0 || device >= n) return fail(RLPT_ERR_ARG, "rlpt_ctx_create: device out of range");
    CK(cudaSetDevice(device));
    rlpt_ctx* c = new rlpt_ctx; c->device = device;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    if (const char* e = getenv("RLPT_SPLIT")) c->pipe_split = atoi(e) != 0;
    if (const char* e = getenv("RLPT_TAIL")) c->pipe_tail = std::max(0, atoi(e));
    if (const char* e = getenv("RLPT_PRE")) c->pipe_pre = atoi(e) != 0;
    if (const char* e = getenv("RLPT_NQ_GRAPH")) c->nq_graphs = atoi(e) != 0;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaMalloc(&c->d_stats, sizeof(unsigned long long) * 8)); CK(cudaMemset(c->d_stats, 0, sizeof(unsigned long long) * 8));
    CK(cudaMalloc(&c->d_cap_n, sizeof(int))); CK(cudaMemset(c->d_cap_n, 0, sizeof(int)));
    CK(cudaMalloc(&c->d_chit_cursor, sizeof(int))); CK(cudaMemset(c->d_chit_cursor, 0, sizeof(int)));
    if (const char* e = getenv("RLPT_BVH_LEAF")) c->bvh_leaf_max = std::max(1, std::min(atoi(e), (int)BVH4_LEAF_MAX));
    rlpt_config_default(&c->cfg);
    float cs[CELLS]; for (int k = 0; k < CELLS; ++k) cs[k] = cell_centre_cos(k);
    upload_cell_cos(cs); dqn_upload_cell_cos(cs);
    int lim = kernels_set_smem_limit((size_t)prop.sharedMemPerBlockOptin - 12288);    // the kernels also hold static shared memory (compaction scratch; k_isect_bvh: leaf lists + per-lane results, 5 KB)
    if (!lim) lim = dqn_set_smem_limit();
    if (lim) { delete c; return fail(RLPT_ERR_CUDA, "rlpt_ctx_create: cudaFuncSetAttribute failed (kernel image missing for this GPU? built for sm_100a only)"); }
    *out = c;
    return RLPT_OK;
}

int rlpt_ctx_destroy(rlpt_ctx* c) {
    if (!c) return RLPT_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_scene(c); free_rmap(c); p2p_destroy(c); free_frame(c); dqn_free(c->dq); dqn_train_free(c->dq_train); cudaFree(c->d_nq_q);
    cudaFree(c->d_stage); cudaFree(c->d_stats); if (c->ev_fork) cudaEventDestroy(c->ev_fork); cudaFree(c->d_cap_o); cudaFree(c->d_cap_d); cudaFree(c->d_cap_n); cudaFree(c->d_chit_cursor);
    for (cudaEvent_t e : c->phase_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->kev_pool) cudaEventDestroy(e);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaStreamDestroy(c->stream);
    delete c;
    return RLPT_OK;
}

int rlpt_sync(rlpt_ctx* c) { if (!c) return fail(RLPT_ERR_ARG, "null ctx"); CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream)); CK(cudaGetLastError()); return p2p_check(c); }
int rlpt_stream(rlpt_ctx* c, void** s) { if (!c || !s) return fail(RLPT_ERR_ARG, "null"); *s = (void*)c->stream; return RLPT_OK; }
int rlpt_config_get(rlpt_ctx* c, rlpt_config* cfg) { if (!c || !cfg) return fail(RLPT_ERR_ARG, "null"); *cfg = c->cfg; return RLPT_OK; }
int rlpt_set_allreduce(rlpt_ctx* c, rlpt_allreduce_fn fn, void* user) { if (!c) return fail(RLPT_ERR_ARG, "null ctx"); c->allreduce = fn; c->allreduce_user = user; return RLPT_OK; }

static int choose_traversal(rlpt_ctx* c, int traversal) {
    int n_tri = c->n_surf + c->n_light;
    size_t optin = 0; { cudaDeviceProp prop; cudaGetDeviceProperties(&prop, c->device); optin = prop.sharedMemPerBlockOptin; }
    SceneDev& sc = c->scene;
    int mode = traversal ? traversal : c->cfg.traversal;
    if (mode == RLPT_TRAVERSAL_AUTO) mode = n_tri <= 48 ? RLPT_TRAVERSAL_BRUTE : RLPT_TRAVERSAL_BVH;
    sc.brute = mode == RLPT_TRAVERSAL_BRUTE;
    // shared-memory budget: the whole scene is staged by every CTA when it is small enough to leave room for four resident CTAs
    // of the closest-hit kernel (BVH scenes: plus its per-warp traversal scratch); otherwise everything comes through the
    // read-only path (L1 hit rate 93-95 % on Medieval_House)
    cudaFree(c->d_scan); cudaFree(c->d_scan_gid); c->d_scan = nullptr; c->d_scan_gid = nullptr;
    sc.scan = nullptr; sc.slot_gid = nullptr; sc.n_units = 0; sc.n_items = 0; sc.bundle = 0; sc.k1 = sc.k2 = sc.k3 = sc.vmax = 0.f;
    sc.staged = 1;
    const size_t scene_b = scene_smem_bytes(sc) + (sc.brute ? (size_t)n_tri * 80 + 64 : 0);      // brute: + scan units and slot table (built below), upper bound
    const size_t budget = optin > 16384 ? optin - 12288 : optin;
    if (sc.brute) {
        if (scene_b > budget) return fail(RLPT_ERR_UNSUPPORTED, "brute-force traversal needs the whole scene in shared memory");
    } else if (scene_b > 20480 || scene_b + bvh4_scratch_bytes() > budget) sc.staged = 0;       // 20 KB: three to four CTAs of k_isect_bvh (50 KB of traversal scratch each) still fit
    c->smem_bytes = scene_smem_bytes(sc);
    {   // |detA| = |a . (e1 x e2)| <= SCREEN_HEIGHT |e1| |e2|: below 2^23 for every primitive -> the guard-free candidate pass applies
        double worst = 0.0;
        auto scan = [&](const std::vector<float>& v) {
            for (size_t g = 0; g + 9 <= v.size(); g += 9) {
                double e1 = 0, e2 = 0;
                for (int k = 0; k < 3; ++k) { double a = (double)v[g + 3 + k] - v[g + k], b = (double)v[g + 6 + k] - v[g + k]; e1 += a * a; e2 += b * b; }
                worst = std::max(worst, std::sqrt(e1) * std::sqrt(e2));
            }
        };
        scan(c->h_surf_v); scan(c->h_light_v);
        sc.det_small = (std::isfinite(worst) && (double)c->cfg.height * worst * 1.001 < 8388608.0) ? 1 : 0;
    }
    // scan units of the conservative pre-test (rlpt_device.cuh, unit_candidates): parallelogram pairs, then single triangles
    if (sc.brute && n_tri <= 64 && !getenv("RLPT_NO_UNITS")) {
        std::vector<float> verts(c->h_surf_v); verts.insert(verts.end(), c->h_light_v.begin(), c->h_light_v.end());
        HostScanUnits hu; host_build_scan_units(verts.data(), n_tri, hu);
        if (hu.n_pairs > 0 && (int)hu.slot_gid.size() == n_tri) {
            while (hu.slot_gid.size() % 4) hu.slot_gid.push_back(0);             // staged four at a time
            CK(cudaMalloc(&c->d_scan, sizeof(float) * hu.scan.size())); CK(cudaMalloc(&c->d_scan_gid, sizeof(int) * hu.slot_gid.size()));
            CK(cudaMemcpy(c->d_scan, hu.scan.data(), sizeof(float) * hu.scan.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(c->d_scan_gid, hu.slot_gid.data(), sizeof(int) * hu.slot_gid.size(), cudaMemcpyHostToDevice));
            sc.scan = reinterpret_cast<const float4*>(c->d_scan); sc.slot_gid = c->d_scan_gid; sc.n_units = hu.n_pairs; sc.n_items = hu.n_items; sc.bundle = getenv("RLPT_NO_BUNDLE") ? 0 : 1;
            sc.k1 = hu.k1; sc.k2 = hu.k2; sc.k3 = hu.k3; sc.vmax = hu.vmax;
            c->smem_bytes = scene_smem_bytes(sc);
        }
    }
    return RLPT_OK;
}

int rlpt_config_set(rlpt_ctx* c, const rlpt_config* cfg) {
    if (!c || !cfg) return fail(RLPT_ERR_ARG, "rlpt_config_set: null");
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->spp <= 0) return fail(RLPT_ERR_ARG, "rlpt_config_set: width, height, spp must be positive");
    if (cfg->max_bounces <= 0 || cfg->max_bounces > 255) return fail(RLPT_ERR_ARG, "rlpt_config_set: max_bounces must be in 1..255");
    if (cfg->world_size <= 0 || cfg->rank < 0 || cfg->rank >= cfg->world_size) return fail(RLPT_ERR_ARG, "rlpt_config_set: bad rank/world_size");
    if ((double)cfg->width * cfg->height * cfg->spp > 2.0e9) return fail(RLPT_ERR_ARG, "rlpt_config_set: width*height*spp exceeds 2^31 paths per frame");
    bool geometry_changed = cfg->width != c->cfg.width || cfg->height != c->cfg.height;
    c->cfg = *cfg;
    CK(cudaSetDevice(c->device));
    if (geometry_changed && c->d_accum) { CK(cudaStreamSynchronize(c->stream)); for (auto& l : c->lanes) if (l.stream) CK(cudaStreamSynchronize(l.stream)); free_frame(c); }
    if (c->have_scene) { int rc = choose_traversal(c, 0); if (rc) return rc; }
    if (c->have_rmap) { c->rm.within_abs = within_abs_of(cfg->max_dist); c->rm.vc.accept_r = std::min(c->vc_built_accept, c->rm.within_abs * (1.f - 1e-5f));   // lists were built for vc_built_accept: valid for any smaller radius
                        c->rm.vx_table = c->rm.within_abs == c->vx_built_within ? c->d_vx_table : nullptr; }          // the second level's lists hold for the search radius they were built with only
    return RLPT_OK;
}

int rlpt_scene_upload(rlpt_ctx* c, const float* sv, const float* srgb, int ns, const float* lv, const float* lrgb, int nl) {
    if (!c) return fail(RLPT_ERR_ARG, "rlpt_scene_upload: null ctx");
    if (ns < 0 || nl < 0 || ns + nl == 0) return fail(RLPT_ERR_ARG, "rlpt_scene_upload: empty scene");
    if ((ns && (!sv || !srgb)) || (nl && (!lv || !lrgb))) return fail(RLPT_ERR_ARG, "rlpt_scene_upload: null array");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    free_scene(c); free_rmap(c);
    const int n = ns + nl;
    c->n_surf = ns; c->n_light = nl;
    c->h_surf_v.assign(sv, sv + 9 * (size_t)ns); c->h_surf_rgb.assign(srgb, srgb + 3 * (size_t)ns);
    c->h_light_v.assign(lv, lv + 9 * (size_t)nl); c->h_light_rgb.assign(lrgb, lrgb + 3 * (size_t)nl);
    c->h_surf_nrm.resize(3 * (size_t)ns); c->h_surf_class.resize(ns);
    std::vector<float4> tri(3 * (size_t)n), shade(4 * (size_t)n); std::vector<float> lum_pi(std::max(ns, 1));
    // normal classes: surfaces whose normals compare equal component-wise (the reference's `normal == leaf.normal`,
    // radiance_map.cu:175) share a class; -0 == +0; a NaN normal equals nothing.
    std::map<std::array<uint32_t, 3>, int> classes;
    for (int g = 0; g < n; ++g) {
        const float* v = g < ns ? sv + 9 * (size_t)g : lv + 9 * (size_t)(g - ns);
        const float* col = g < ns ? srgb + 3 * (size_t)g : lrgb + 3 * (size_t)(g - ns);
        float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
        float T1 = fmaf(e1[1], e2[2], -(e1[2] * e2[1]));
        tri[3 * g] = make_float4(v[0], v[1], v[2], e1[0]); tri[3 * g + 1] = make_float4(e1[1], e1[2], e2[0], e2[1]); tri[3 * g + 2] = make_float4(e2[2], T1, 0.f, 0.f);
        float nrm[3]; host_triangle_normal(v, nrm);
        f3 N = { nrm[0], nrm[1], nrm[2] }, T, B; tangent_frame(N, T, B);
        float lum = luminance3(col);
        int cls;
        if (std::isnan(nrm[0]) || std::isnan(nrm[1]) || std::isnan(nrm[2])) cls = -2;
        else {
            std::array<uint32_t, 3> key; for (int k = 0; k < 3; ++k) { float f = nrm[k] == 0.f ? 0.f : nrm[k]; memcpy(&key[k], &f, 4); }
            auto it = classes.find(key); if (it == classes.end()) it = classes.emplace(key, (int)classes.size()).first; cls = it->second;
        }
        if (g < ns) {
            for (int k = 0; k < 3; ++k) c->h_surf_nrm[3 * (size_t)g + k] = nrm[k];
            c->h_surf_class[g] = cls; lum_pi[g] = lum / PI_F;
            shade[4 * g] = make_float4(nrm[0], nrm[1], nrm[2], lum / PI_F);
            shade[4 * g + 3] = make_float4(col[0] / PI_F, col[1] / PI_F, col[2] / PI_F, lum);     // BRDF = diffuse_c / (float)M_PI; w = luminance (Neural-Q discount)
        } else {
            shade[4 * g] = make_float4(nrm[0], nrm[1], nrm[2], lum);
            shade[4 * g + 3] = make_float4(col[0], col[1], col[2], 0.f);
        }
        int qcls = cls == -2 ? -3 : cls;           // a query with a NaN normal matches no leaf
        float fc; memcpy(&fc, &qcls, 4);
        shade[4 * g + 1] = make_float4(T.x, T.y, T.z, fc);
        shade[4 * g + 2] = make_float4(B.x, B.y, B.z, 0.f);
    }
    CK(cudaMalloc(&c->d_tri, sizeof(float4) * tri.size())); CK(cudaMalloc(&c->d_shade, sizeof(float4) * shade.size()));
    CK(cudaMalloc(&c->d_surf_lum_over_pi, sizeof(float) * lum_pi.size()));
    CK(cudaMemcpyAsync(c->d_tri, tri.data(), sizeof(float4) * tri.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_shade, shade.data(), sizeof(float4) * shade.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_surf_lum_over_pi, lum_pi.data(), sizeof(float) * lum_pi.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int n_nodes = 0, depth = 0, n_nodes4 = 0, depth4 = 0;
    int rc = bvh_build_gpu(c->d_tri, n, c->bvh_leaf_max, &c->d_bvh, &n_nodes, &depth, &c->d_bvh4, &n_nodes4, &depth4, &c->d_tri4, c->stream);
    if (rc) { char b[128]; snprintf(b, sizeof b, "rlpt_scene_upload: GPU BVH build failed (%d: %s)", rc, rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "topology"); return fail(RLPT_ERR_CUDA, b); }
    if (depth > 30) return fail(RLPT_ERR_UNSUPPORTED, "rlpt_scene_upload: BVH deeper than the traversal stack");
    c->bvh_depth = depth; c->bvh4_depth = depth4; c->n_nodes = n_nodes;
    c->scene = SceneDev{}; c->scene.tri = c->d_tri; c->scene.shade = c->d_shade; c->scene.bvh4 = c->d_bvh4; c->scene.tri4 = c->d_tri4;
    c->scene.n_tri = n; c->scene.n_surf = ns; c->scene.n_light = nl; c->scene.n_nodes4 = n_nodes4;
    c->have_scene = true;
    c->dq_vertices.clear(); c->dq_vertices_custom = false; c->dq.ready = false;
    return choose_traversal(c, 0);
}

int rlpt_scene_info(rlpt_ctx* c, int* ns, int* nl, int* nodes, int* depth) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_scene_info: no scene");
    if (ns) *ns = c->n_surf; if (nl) *nl = c->n_light; if (nodes) *nodes = c->n_nodes; if (depth) *depth = c->bvh_depth;
    return RLPT_OK;
}
int rlpt_scene_bvh_download(rlpt_ctx* c, float* nodes16, int max_nodes) {
    if (!c || !c->have_scene || !nodes16) return fail(RLPT_ERR_ARG, "rlpt_scene_bvh_download: no scene");
    CK(cudaSetDevice(c->device));
    int n = std::min(max_nodes, c->n_nodes);
    CK(cudaMemcpy(nodes16, c->d_bvh, sizeof(float) * 16 * (size_t)n, cudaMemcpyDeviceToHost));
    return RLPT_OK;
}
// the 4-wide tree the kernels walk (rlpt_bvh.cu): 28 floats per node, and the primitive id of every leaf-order triangle record
int rlpt_scene_bvh4_info(rlpt_ctx* c, int* nodes, int* depth, int* leaf_max) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_scene_bvh4_info: no scene");
    if (nodes) *nodes = c->scene.n_nodes4; if (depth) *depth = c->bvh4_depth; if (leaf_max) *leaf_max = c->bvh_leaf_max;
    return RLPT_OK;
}
int rlpt_scene_bvh4_download(rlpt_ctx* c, float* nodes28, int max_nodes, int* record_gid, int max_records) {
    if (!c || !c->have_scene || !nodes28) return fail(RLPT_ERR_ARG, "rlpt_scene_bvh4_download: no scene");
    CK(cudaSetDevice(c->device));
    const int n = std::min(max_nodes, c->scene.n_nodes4);
    CK(cudaMemcpy(nodes28, c->d_bvh4, sizeof(float) * 28 * (size_t)n, cudaMemcpyDeviceToHost));
    if (record_gid) {
        const int m = std::min(max_records, c->scene.n_tri);
        std::vector<float4> rec(3 * (size_t)m);
        CK(cudaMemcpy(rec.data(), c->d_tri4, sizeof(float4) * rec.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; ++i) memcpy(&record_gid[i], &rec[3 * (size_t)i + 2].z, 4);
    }
    return RLPT_OK;
}

int rlpt_camera_set(rlpt_ctx* c, const float position[4], float yaw_y, float yaw_x) {
    if (!c || !position) return fail(RLPT_ERR_ARG, "rlpt_camera_set: null");
    c->cam[0] = position[0]; c->cam[1] = position[1]; c->cam[2] = position[2]; c->yaw_y = yaw_y; c->yaw_x = yaw_x;
    return RLPT_OK;
}

// replaces: the greedy samplers RadianceVolume::sample_max_direction_from_radiance_distribution (G/radiance_volumes/radiance_volume.cu:248-278) and
// sample_max_direction (G/deep_learning/nn_rendering_helpers.cu:492-553), which the reference swaps in by hand for its 1-spp "what has been learned" figures
int rlpt_set_max_direction(rlpt_ctx* c, int on) { if (!c) return fail(RLPT_ERR_ARG, "null ctx"); c->max_dir = on ? 1 : 0; return RLPT_OK; }

int rlpt_closest_hit_device(rlpt_ctx* c, const float* d_org, const float* d_dir, int n, int traversal, int* d_type, int* d_index, float* d_t, unsigned long long* d_counters) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_closest_hit: upload a scene first");
    if (n < 0) return fail(RLPT_ERR_ARG, "rlpt_closest_hit: negative ray count");
    if (n == 0) return RLPT_OK;
    CK(cudaSetDevice(c->device));
    // a traversal override lasts for this call: choose_traversal owns device buffers (scan units), so the configured mode
    // is re-established by calling it again rather than by restoring a saved copy of the scene descriptor
    if (traversal) { int rc = choose_traversal(c, traversal); if (rc) { (void)choose_traversal(c, 0); return rc; } }
    launch_closest_hit(c->scene, d_org, d_dir, n, (float)c->cfg.height, d_type, d_index, d_t, d_counters, c->d_chit_cursor, c->smem_bytes, c->stream);
    if (traversal) { CK(cudaStreamSynchronize(c->stream)); int rc = choose_traversal(c, 0); if (rc) return rc; }
    CK(cudaGetLastError());
    return RLPT_OK;
}

int rlpt_closest_hit(rlpt_ctx* c, const float* org, const float* dir, int n, int traversal, int* type, int* index, float* t, unsigned long long* counters) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_closest_hit: upload a scene first");
    if (n < 0) return fail(RLPT_ERR_ARG, "rlpt_closest_hit: negative ray count");
    if (n == 0) return RLPT_OK;
    if (!org || !dir || !type || !index || !t) return fail(RLPT_ERR_ARG, "rlpt_closest_hit: null array");
    CK(cudaSetDevice(c->device));
    float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr; int *d_ty = nullptr, *d_ix = nullptr; unsigned long long* d_c = nullptr;
    CK(cudaMalloc(&d_o, sizeof(float) * 3 * (size_t)n)); CK(cudaMalloc(&d_d, sizeof(float) * 3 * (size_t)n)); CK(cudaMalloc(&d_t, sizeof(float) * (size_t)n));
    CK(cudaMalloc(&d_ty, sizeof(int) * (size_t)n)); CK(cudaMalloc(&d_ix, sizeof(int) * (size_t)n));
    if (counters) { CK(cudaMalloc(&d_c, sizeof(unsigned long long) * 2)); CK(cudaMemsetAsync(d_c, 0, sizeof(unsigned long long) * 2, c->stream)); }
    CK(cudaMemcpyAsync(d_o, org, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_d, dir, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    int rc = rlpt_closest_hit_device(c, d_o, d_d, n, traversal, d_ty, d_ix, d_t, d_c);
    if (rc == RLPT_OK) {
        CK(cudaMemcpyAsync(type, d_ty, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(index, d_ix, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(t, d_t, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        if (counters) CK(cudaMemcpyAsync(counters, d_c, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_ty); cudaFree(d_ix); cudaFree(d_c);
    return rc;
}

// ------------------------------------------------------------------------------------------------ radiance map
int rlpt_radiance_map_build(rlpt_ctx* c) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_build: upload a scene first");
    if (c->n_surf == 0) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_build: scene has no surfaces");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    free_rmap(c);
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    host_build_radiance_map(c->h_surf_v.data(), c->h_surf_nrm.data(), c->n_surf, c->cfg.area_per_sample, c->h_vol, c->h_tree);
    c->rmap_build_s[0] = since(t_begin); c->rmap_build_s[1] = 0.0;
    const int nv = (int)c->h_vol.size(), nt = (int)c->h_tree.size();
    if (nv == 0) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_build: no radiance volumes (surfaces smaller than area_per_sample)");
    if (nv >= (1 << 24)) return fail(RLPT_ERR_UNSUPPORTED, "rlpt_radiance_map_build: more than 2^24 radiance volumes");
    // re-lay the reference's flattened tree out as inner nodes only; leaf children carry the volume in the child word
    std::vector<int> inner_of(nt, -1); int n_inner = 0;
    for (int i = 0; i < nt; ++i) if (!c->h_tree[i].leaf) inner_of[i] = n_inner++;
    auto child_word = [&](unsigned idx) -> uint32_t { const HostTreeElement& e = c->h_tree[idx]; return e.leaf ? (KD_LEAF | (uint32_t)(int)e.data) : (uint32_t)inner_of[idx]; };
    std::vector<float4> kd(std::max(n_inner, 1));
    for (int i = 0; i < nt; ++i) {
        const HostTreeElement& e = c->h_tree[i]; if (e.leaf) continue;
        uint32_t l = child_word(e.left), r = child_word(e.right); float fl, fr, fd; int dim = e.dim;
        memcpy(&fl, &l, 4); memcpy(&fr, &r, 4); memcpy(&fd, &dim, 4);
        kd[inner_of[i]] = make_float4(e.data, fl, fr, fd);
    }
    std::vector<float4> posn(nv); std::vector<int> vsurf(nv);
    for (int i = 0; i < nv; ++i) {
        int cls = c->h_surf_class[c->h_vol[i].surface]; float fc; memcpy(&fc, &cls, 4);
        posn[i] = make_float4(c->h_vol[i].pos[0], c->h_vol[i].pos[1], c->h_vol[i].pos[2], fc); vsurf[i] = c->h_vol[i].surface;
    }
    const size_t cells = (size_t)nv * CELLS;
    CK(cudaMalloc(&c->d_kd, sizeof(float4) * kd.size())); CK(cudaMalloc(&c->d_posn, sizeof(float4) * nv)); CK(cudaMalloc(&c->d_vol_surface, sizeof(int) * nv));
    CK(cudaMalloc(&c->d_q, 4 * cells)); CK(cudaMalloc(&c->d_cdf, 4 * cells)); CK(cudaMalloc(&c->d_cdf_rows, 4 * (size_t)nv * GRID)); CK(cudaMalloc(&c->d_visits, 4 * cells)); CK(cudaMalloc(&c->d_irr, 4 * (size_t)nv));
    CK(cudaMalloc(&c->d_acc_sum, 4 * cells)); CK(cudaMalloc(&c->d_acc_cnt, 4 * cells));
    CK(cudaMemcpy(c->d_kd, kd.data(), sizeof(float4) * kd.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_posn, posn.data(), sizeof(float4) * nv, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_vol_surface, vsurf.data(), sizeof(int) * nv, cudaMemcpyHostToDevice));
    // initial state: Q = INITIAL_RADIANCE everywhere, visits 0 (radiance_volume.cu:49-89); the CDF and the irradiance
    // estimate are derived from Q by the merge kernel (deliberate deviation: the reference starts from the exclusive
    // CDF k/144, under which 1 sample in 144 fails until the first update)
    std::vector<float> q0(cells, c->cfg.initial_radiance);
    CK(cudaMemcpy(c->d_q, q0.data(), 4 * cells, cudaMemcpyHostToDevice));
    CK(cudaMemset(c->d_visits, 0, 4 * cells)); CK(cudaMemset(c->d_acc_sum, 0, 4 * cells)); CK(cudaMemset(c->d_acc_cnt, 0, 4 * cells));
    RadianceDev& rm = c->rm;
    rm.kd_inner = c->d_kd; rm.vol_posn = c->d_posn; rm.vol_surface = c->d_vol_surface; rm.q = c->d_q; rm.cdf = c->d_cdf; rm.cdf_rows = c->d_cdf_rows; rm.visits = c->d_visits;
    rm.irradiance = c->d_irr; rm.acc_sum = c->d_acc_sum; rm.acc_cnt = c->d_acc_cnt; rm.n_vol = nv; rm.n_inner = n_inner;
    rm.root = child_word(0);
    rm.root_px = c->h_tree[0].pos[0]; rm.root_py = c->h_tree[0].pos[1]; rm.root_pz = c->h_tree[0].pos[2];
    rm.within_abs = within_abs_of(c->cfg.max_dist);
    {
        // candidate cells for the nearest-volume search: cell size = a fraction of the volume spacing sqrt(AREA_PER_SAMPLE)
        float factor = 0.6f; if (const char* e = getenv("RLPT_VCELL")) factor = std::max(0.05f, (float)atof(e));
        const float accept = rm.within_abs * (1.f - 1e-5f);
        std::vector<int> vclass(nv); for (int i = 0; i < nv; ++i) vclass[i] = c->h_surf_class[c->h_vol[i].surface];
        HostVCells hv;
        const auto t_vc = std::chrono::steady_clock::now();
        host_build_vcells(c->h_surf_v.data(), c->h_surf_class.data(), c->n_surf, c->h_vol, vclass, c->h_tree, factor * std::sqrt(std::max(c->cfg.area_per_sample, 1e-12f)), accept, rm.within_abs, hv);
        c->rmap_build_s[1] = since(t_vc);
        CK(cudaMalloc(&c->d_vc_table, sizeof(int) * hv.table.size())); CK(cudaMalloc(&c->d_vc_cand, sizeof(float) * hv.cand.size()));
        CK(cudaMemcpy(c->d_vc_table, hv.table.data(), sizeof(int) * hv.table.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_vc_cand, hv.cand.data(), sizeof(float) * hv.cand.size(), cudaMemcpyHostToDevice));
        VCells g{}; g.ox = hv.ox; g.oy = hv.oy; g.oz = hv.oz; g.inv_h = 1.f / hv.h; g.nx = hv.nx; g.ny = hv.ny; g.nz = hv.nz;
        g.mask = (uint32_t)(hv.table.size() / 4 - 1); g.accept_r = accept; c->vc_built_accept = accept;
        c->vc_keys = hv.keys; c->vc_listed = hv.listed;
        rm.vc = g; rm.vc_table = c->d_vc_table; rm.vc_cand = c->d_vc_cand;
        CK(cudaMalloc(&c->d_vx_table, sizeof(int) * hv.xtable.size())); CK(cudaMalloc(&c->d_vx_cand, sizeof(float) * hv.xcand.size()));
        CK(cudaMemcpy(c->d_vx_table, hv.xtable.data(), sizeof(int) * hv.xtable.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_vx_cand, hv.xcand.data(), sizeof(float) * hv.xcand.size(), cudaMemcpyHostToDevice));
        rm.vx_table = c->d_vx_table; rm.vx_cand = c->d_vx_cand; rm.vx_mask = (uint32_t)(hv.xtable.size() / 4 - 1); c->vx_built_within = rm.within_abs;
    }
    c->have_rmap = true;
    launch_merge(rm, c->d_surf_lum_over_pi, c->cfg.radiance_threshold, 1, c->stream);
    CK(cudaStreamSynchronize(c->stream)); CK(cudaGetLastError());
    c->rmap_build_s[3] = since(t_begin); c->rmap_build_s[2] = c->rmap_build_s[3] - c->rmap_build_s[0] - c->rmap_build_s[1];
    return RLPT_OK;
}
int rlpt_radiance_map_build_seconds(rlpt_ctx* c, double* seconds4) {
    if (!c || !c->have_rmap || !seconds4) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_build_seconds: no radiance map");
    for (int k = 0; k < 4; ++k) seconds4[k] = c->rmap_build_s[k];
    return RLPT_OK;
}

// ---- peer-memory exchange setup. Blob: n_volumes (int, padded to 8 bytes) + 8 cudaIpcMemHandle_t
// (acc_sum, acc_cnt, q, cdf, cdf_rows, visits, irradiance, flags).
static const int P2P_HANDLES = 8;
int rlpt_p2p_blob_bytes(void) { return 8 + P2P_HANDLES * (int)sizeof(cudaIpcMemHandle_t); }
int rlpt_p2p_export(rlpt_ctx* c, void* blob) {
    if (!c || !c->have_rmap || !blob) return fail(RLPT_ERR_ARG, "rlpt_p2p_export: build the radiance map first");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    if (!c->d_p2p_flags) {
        CK(cudaMalloc(&c->d_p2p_flags, sizeof(unsigned) * 2 * MAX_PEERS)); CK(cudaMemset(c->d_p2p_flags, 0, sizeof(unsigned) * 2 * MAX_PEERS));
        CK(cudaMalloc(&c->d_p2p_done, sizeof(unsigned))); CK(cudaMemset(c->d_p2p_done, 0, sizeof(unsigned)));
        CK(cudaMalloc(&c->d_p2p_error, sizeof(unsigned))); CK(cudaMemset(c->d_p2p_error, 0, sizeof(unsigned)));
    }
    unsigned char* out = (unsigned char*)blob; memset(out, 0, 8);
    const int nv = c->rm.n_vol; memcpy(out, &nv, 4); memcpy(out + 4, &c->p2p_epoch, 4);        // the epoch this rank has reached: import resumes from the largest
    void* ptrs[P2P_HANDLES] = { c->d_acc_sum, c->d_acc_cnt, c->d_q, c->d_cdf, c->d_cdf_rows, c->d_visits, c->d_irr, c->d_p2p_flags };
    for (int i = 0; i < P2P_HANDLES; ++i) { cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, ptrs[i])); memcpy(out + 8 + i * sizeof h, &h, sizeof h); }
    return RLPT_OK;
}
int rlpt_p2p_import(rlpt_ctx* c, const void* blobs, int world_size) {
    if (!c || !c->have_rmap || !blobs || !c->d_p2p_flags) return fail(RLPT_ERR_ARG, "rlpt_p2p_import: call rlpt_p2p_export first");
    if (world_size != c->cfg.world_size || world_size < 2 || world_size > MAX_PEERS) return fail(RLPT_ERR_ARG, "rlpt_p2p_import: world size must equal rlpt_config.world_size and be 2..8");
    CK(cudaSetDevice(c->device));
    for (void* p : c->p2p_opened) cudaIpcCloseMemHandle(p);
    c->p2p_opened.clear(); c->p2p_ready = false;
    PeerTables pt{}; pt.world = world_size; pt.rank = c->cfg.rank; pt.error = c->d_p2p_error;
    const size_t stride = (size_t)rlpt_p2p_blob_bytes();
    unsigned epoch = c->p2p_epoch;
    for (int r = 0; r < world_size; ++r) {
        const unsigned char* in = (const unsigned char*)blobs + (size_t)r * stride;
        int nv = 0; memcpy(&nv, in, 4);
        unsigned er = 0; memcpy(&er, in + 4, 4); if ((int)(er - epoch) > 0) epoch = er;
        if (nv != c->rm.n_vol) return fail(RLPT_ERR_ARG, "rlpt_p2p_import: rank " + std::to_string(r) + " has a different radiance map");
        void* ptrs[P2P_HANDLES];
        if (r == c->cfg.rank) {
            void* own[P2P_HANDLES] = { c->d_acc_sum, c->d_acc_cnt, c->d_q, c->d_cdf, c->d_cdf_rows, c->d_visits, c->d_irr, c->d_p2p_flags };
            memcpy(ptrs, own, sizeof own);
        } else {
            for (int i = 0; i < P2P_HANDLES; ++i) {
                cudaIpcMemHandle_t h; memcpy(&h, in + 8 + i * sizeof h, sizeof h);
                cudaError_t e = cudaIpcOpenMemHandle(&ptrs[i], h, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) return fail(RLPT_ERR_CUDA, std::string("rlpt_p2p_import: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
                c->p2p_opened.push_back(ptrs[i]);
            }
        }
        pt.acc_sum[r] = (float*)ptrs[0]; pt.acc_cnt[r] = (uint32_t*)ptrs[1]; pt.q[r] = (float*)ptrs[2]; pt.cdf[r] = (float*)ptrs[3]; pt.cdf_rows[r] = (float*)ptrs[4];
        pt.visits[r] = (uint32_t*)ptrs[5]; pt.irradiance[r] = (float*)ptrs[6]; pt.flags[r] = (unsigned*)ptrs[7];
    }
    c->peers = pt; c->p2p_ready = true; c->p2p_epoch = epoch;         // every rank continues from the same epoch
    CK(cudaMemset(c->d_p2p_error, 0, sizeof(unsigned))); CK(cudaMemset(c->d_p2p_done, 0, sizeof(unsigned)));
    return RLPT_OK;
}

int rlpt_p2p_close(rlpt_ctx* c) {
    if (!c) return fail(RLPT_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    for (void* p : c->p2p_opened) cudaIpcCloseMemHandle(p);
    c->p2p_opened.clear(); c->p2p_ready = false;                 // the flag arrays stay: peers may still hold them open
    return RLPT_OK;
}

int rlpt_radiance_map_info(rlpt_ctx* c, int* nv, int* nt) {
    if (!c || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_info: no radiance map");
    if (nv) *nv = (int)c->h_vol.size(); if (nt) *nt = (int)c->h_tree.size();
    return RLPT_OK;
}
int rlpt_radiance_map_tree(rlpt_ctx* c, int* dim, int* leaf, unsigned* left, unsigned* right, float* data, float* pos3, float* nrm3) {
    if (!c || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_tree: no radiance map");
    for (size_t i = 0; i < c->h_tree.size(); ++i) {
        const HostTreeElement& e = c->h_tree[i];
        if (dim) dim[i] = e.dim; if (leaf) leaf[i] = e.leaf; if (left) left[i] = e.left; if (right) right[i] = e.right; if (data) data[i] = e.data;
        for (int k = 0; k < 3; ++k) { if (pos3) pos3[3 * i + k] = e.pos[k]; if (nrm3) nrm3[3 * i + k] = e.nrm[k]; }
    }
    return RLPT_OK;
}

int rlpt_radiance_map_find_closest(rlpt_ctx* c, const float* pos, const float* nrm, int n, int* out) {
    if (!c || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_find_closest: no radiance map");
    if (n <= 0) return n == 0 ? RLPT_OK : fail(RLPT_ERR_ARG, "negative count");
    CK(cudaSetDevice(c->device));
    // query normal -> normal class (exact component-wise equality with some surface's normal, else "matches nothing")
    std::vector<int> cls(n, -3);
    for (int i = 0; i < n; ++i) {
        const float* q = nrm + 3 * (size_t)i;
        for (int s = 0; s < c->n_surf; ++s) { const float* m = &c->h_surf_nrm[3 * (size_t)s]; if (q[0] == m[0] && q[1] == m[1] && q[2] == m[2]) { cls[i] = c->h_surf_class[s]; break; } }
    }
    float* d_pos = nullptr; int *d_cls = nullptr, *d_out = nullptr;
    CK(cudaMalloc(&d_pos, sizeof(float) * 3 * (size_t)n)); CK(cudaMalloc(&d_cls, sizeof(int) * (size_t)n)); CK(cudaMalloc(&d_out, sizeof(int) * (size_t)n));
    CK(cudaMemcpyAsync(d_pos, pos, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_cls, cls.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    launch_find_closest(c->rm, c->scene, d_pos, reinterpret_cast<const float*>(d_cls), n, d_out, c->stream);
    CK(cudaMemcpyAsync(out, d_out, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream)); CK(cudaGetLastError());
    cudaFree(d_pos); cudaFree(d_cls); cudaFree(d_out);
    return RLPT_OK;
}

int rlpt_radiance_map_set_q(rlpt_ctx* c, const float* q, const unsigned* visits) {
    if (!c || !c->have_rmap || !q) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_set_q: no radiance map / null");
    CK(cudaSetDevice(c->device));
    const size_t cells = (size_t)c->rm.n_vol * CELLS;
    CK(cudaMemcpyAsync(c->d_q, q, 4 * cells, cudaMemcpyHostToDevice, c->stream));
    if (visits) CK(cudaMemcpyAsync(c->d_visits, visits, 4 * cells, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RLPT_OK;
}
int rlpt_radiance_map_update_distributions(rlpt_ctx* c) {
    if (!c || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_update_distributions: no radiance map");
    CK(cudaSetDevice(c->device));
    launch_merge(c->rm, c->d_surf_lum_over_pi, c->cfg.radiance_threshold, 1, c->stream);
    CK(cudaGetLastError());
    return RLPT_OK;
}
int rlpt_radiance_map_download(rlpt_ctx* c, float* q, float* cdf, unsigned* visits, float* irr, float* pos3, float* nrm3, int* surface) {
    if (!c || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_download: no radiance map");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    const size_t cells = (size_t)c->rm.n_vol * CELLS;
    if (q) CK(cudaMemcpy(q, c->d_q, 4 * cells, cudaMemcpyDeviceToHost));
    if (cdf) CK(cudaMemcpy(cdf, c->d_cdf, 4 * cells, cudaMemcpyDeviceToHost));
    if (visits) CK(cudaMemcpy(visits, c->d_visits, 4 * cells, cudaMemcpyDeviceToHost));
    if (irr) CK(cudaMemcpy(irr, c->d_irr, 4 * (size_t)c->rm.n_vol, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < c->h_vol.size(); ++i) {
        for (int k = 0; k < 3; ++k) { if (pos3) pos3[3 * i + k] = c->h_vol[i].pos[k]; if (nrm3) nrm3[3 * i + k] = c->h_vol[i].nrm[k]; }
        if (surface) surface[i] = c->h_vol[i].surface;
    }
    return RLPT_OK;
}
int rlpt_radiance_map_delta_download(rlpt_ctx* c, float* sum, unsigned* cnt) {
    if (!c || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_delta_download: no radiance map");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    const size_t cells = (size_t)c->rm.n_vol * CELLS;
    if (sum) CK(cudaMemcpy(sum, c->d_acc_sum, 4 * cells, cudaMemcpyDeviceToHost));
    if (cnt) CK(cudaMemcpy(cnt, c->d_acc_cnt, 4 * cells, cudaMemcpyDeviceToHost));
    return RLPT_OK;
}

int rlpt_radiance_map_save_q(rlpt_ctx* c, const char* path) {
    if (!c || !c->have_rmap || !path) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_save_q: no radiance map / null path");
    const int nv = c->rm.n_vol; std::vector<float> q((size_t)nv * CELLS);
    int rc = rlpt_radiance_map_download(c, q.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); if (rc) return rc;
    FILE* f = fopen(path, "w");
    if (!f) return fail(RLPT_ERR_IO, std::string("Unable to save the RadianceMap: ") + path);
    fprintf(f, "%d\n", CELLS);
    for (int i = 0; i < nv; ++i) {
        fprintf(f, "%g %g %g", c->h_vol[i].pos[0], c->h_vol[i].pos[1], c->h_vol[i].pos[2]);     // ofstream << float prints %g with 6 digits
        for (int k = 0; k < CELLS; ++k) fprintf(f, " %g", q[(size_t)i * CELLS + k]);
        fprintf(f, "\n");
    }
    fclose(f);
    return RLPT_OK;
}
int rlpt_radiance_map_load_q(rlpt_ctx* c, const char* path) {
    if (!c || !c->have_rmap || !path) return fail(RLPT_ERR_ARG, "rlpt_radiance_map_load_q: no radiance map / null path");
    FILE* f = fopen(path, "r");
    if (!f) return fail(RLPT_ERR_IO, std::string("Could not read radiance volumes: ") + path);
    int cells = 0; const int nv = c->rm.n_vol;
    if (fscanf(f, "%d", &cells) != 1 || cells != CELLS) { fclose(f); return fail(RLPT_ERR_IO, "rlpt_radiance_map_load_q: first line must be 144"); }
    std::vector<float> q((size_t)nv * CELLS);
    for (int i = 0; i < nv; ++i) {
        float p[3]; if (fscanf(f, "%f %f %f", p, p + 1, p + 2) != 3) { fclose(f); return fail(RLPT_ERR_IO, "rlpt_radiance_map_load_q: file has fewer volumes than the map"); }
        for (int k = 0; k < CELLS; ++k) if (fscanf(f, "%f", &q[(size_t)i * CELLS + k]) != 1) { fclose(f); return fail(RLPT_ERR_IO, "rlpt_radiance_map_load_q: truncated row"); }
    }
    fclose(f);
    int rc = rlpt_radiance_map_set_q(c, q.data(), nullptr); if (rc) return rc;
    return rlpt_radiance_map_update_distributions(c);
}

// ------------------------------------------------------------------------------------------------ Neural-Q network
static const std::vector<float>& dqn_vertices(rlpt_ctx* c) {
    if (c->dq_vertices.empty()) { c->dq_vertices = c->h_surf_v; c->dq_vertices.insert(c->dq_vertices.end(), c->h_light_v.begin(), c->h_light_v.end()); }
    return c->dq_vertices;
}
static int dqn_push(rlpt_ctx* c) {          // host parameters -> device, operands derived
    const std::vector<float>& v = dqn_vertices(c);
    if ((int)v.size() != c->dq_host.k_in) return fail(RLPT_ERR_ARG, "DQN input width " + std::to_string(c->dq_host.k_in) + " does not match the scene (" + std::to_string(v.size()) + " vertex floats)");
    nq_graph_reset(c); dqn_train_free(c->dq_train);  // new parameters: fresh optimiser state (and the captured step points into the old buffers)
    int rc = dqn_upload(c->dq, c->dq_host, v.data(), c->stream);
    if (rc) return fail(RLPT_ERR_CUDA, std::string("DQN upload failed: ") + cudaGetErrorString((cudaError_t)rc));
    return RLPT_OK;
}
int rlpt_dqn_set_vertices(rlpt_ctx* c, const float* vertices, int count) {
    if (!c || !c->have_scene || !vertices) return fail(RLPT_ERR_ARG, "rlpt_dqn_set_vertices: upload a scene first");
    if (count != 9 * (c->n_surf + c->n_light)) return fail(RLPT_ERR_ARG, "rlpt_dqn_set_vertices: expected 9 floats per triangle of the uploaded scene");
    c->dq_vertices.assign(vertices, vertices + count); c->dq_vertices_custom = true; c->dq.ready = false;
    return RLPT_OK;
}
int rlpt_dqn_init(rlpt_ctx* c, uint32_t seed) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_dqn_init: upload a scene first");
    CK(cudaSetDevice(c->device));
    dqn_init_glorot(c->dq_host, (int)dqn_vertices(c).size(), seed);
    return dqn_push(c);
}
int rlpt_dqn_load_text(rlpt_ctx* c, const char* path) {
    if (!c || !c->have_scene || !path) return fail(RLPT_ERR_ARG, "rlpt_dqn_load_text: upload a scene first");
    CK(cudaSetDevice(c->device));
    std::string err; if (dqn_load_text(c->dq_host, path, err)) return fail(RLPT_ERR_IO, "rlpt_dqn_load_text: " + err);
    return dqn_push(c);
}
int rlpt_dqn_save_text(rlpt_ctx* c, const char* path) {
    if (!c || !c->dq.ready || !path) return fail(RLPT_ERR_ARG, "rlpt_dqn_save_text: no network");
    CK(cudaSetDevice(c->device));
    int rc = dqn_download(c->dq, c->dq_host, c->stream); if (rc) return fail(RLPT_ERR_CUDA, "DQN download failed");
    std::string err; if (dqn_save_text(c->dq_host, path, err)) return fail(RLPT_ERR_IO, "rlpt_dqn_save_text: " + err);
    return RLPT_OK;
}
static int dqn_total(const DqnHost& h) { int n = 0; for (int l = 0; l < 4; ++l) n += DqnHost::rows(l) * h.cols(l) + DqnHost::rows(l); return n; }
int rlpt_dqn_param_count(rlpt_ctx* c, int* count, int* k_in) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_dqn_param_count: upload a scene first");
    DqnHost shape; shape.k_in = (int)dqn_vertices(c).size();
    if (count) *count = dqn_total(shape); if (k_in) *k_in = shape.k_in;
    return RLPT_OK;
}
int rlpt_dqn_set_params(rlpt_ctx* c, const float* params, int count) {
    if (!c || !c->have_scene || !params) return fail(RLPT_ERR_ARG, "rlpt_dqn_set_params: upload a scene first");
    CK(cudaSetDevice(c->device));
    DqnHost& h = c->dq_host; h.k_in = (int)dqn_vertices(c).size();
    if (count != dqn_total(h)) return fail(RLPT_ERR_ARG, "rlpt_dqn_set_params: wrong parameter count");
    const float* p = params;
    for (int l = 0; l < 4; ++l) { size_t nw = (size_t)DqnHost::rows(l) * h.cols(l); h.w[l].assign(p, p + nw); p += nw; h.b[l].assign(p, p + DqnHost::rows(l)); p += DqnHost::rows(l); }
    return dqn_push(c);
}
int rlpt_dqn_get_params(rlpt_ctx* c, float* params, int count) {
    if (!c || !c->dq.ready || !params) return fail(RLPT_ERR_ARG, "rlpt_dqn_get_params: no network");
    CK(cudaSetDevice(c->device));
    if (dqn_download(c->dq, c->dq_host, c->stream)) return fail(RLPT_ERR_CUDA, "DQN download failed");
    const DqnHost& h = c->dq_host;
    if (count != dqn_total(h)) return fail(RLPT_ERR_ARG, "rlpt_dqn_get_params: wrong parameter count");
    float* p = params;
    for (int l = 0; l < 4; ++l) { p = std::copy(h.w[l].begin(), h.w[l].end(), p); p = std::copy(h.b[l].begin(), h.b[l].end(), p); }
    return RLPT_OK;
}
int rlpt_dqn_forward(rlpt_ctx* c, const float* pos3, int n, float* q) {
    if (!c || !c->dq.ready) return fail(RLPT_ERR_ARG, "rlpt_dqn_forward: no network (rlpt_dqn_init / rlpt_dqn_load_text first)");
    if (n < 0 || (n && (!pos3 || !q))) return fail(RLPT_ERR_ARG, "rlpt_dqn_forward: bad arguments");
    if (n == 0) return RLPT_OK;
    CK(cudaSetDevice(c->device));
    std::vector<float4> h_pos(n); for (int i = 0; i < n; ++i) h_pos[i] = make_float4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], 0.f);
    float4* d_pos = nullptr; float* d_q = nullptr;
    CK(cudaMalloc(&d_pos, sizeof(float4) * (size_t)n)); CK(cudaMalloc(&d_q, sizeof(float) * (size_t)n * DQ_OUT));
    CK(cudaMemcpyAsync(d_pos, h_pos.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    DqnFwdParams p{}; p.pos = d_pos; p.n = n; p.c1 = c->dq.c1; p.m1 = c->dq.m1; p.b2 = c->dq.b[1]; p.b3 = c->dq.b[2]; p.b4 = c->dq.b[3];
    p.w2p = c->dq.w2p; p.w3p = c->dq.w3p; p.w4p = c->dq.w4p; p.q = d_q; p.q_stride = n;
    int rc = dqn_forward(c->dq, p, c->stream);
    if (rc) { cudaFree(d_pos); cudaFree(d_q); return fail(RLPT_ERR_CUDA, std::string("rlpt_dqn_forward: launch failed: ") + cudaGetErrorString((cudaError_t)rc)); }
    std::vector<float> qt((size_t)n * DQ_OUT);
    CK(cudaMemcpyAsync(qt.data(), d_q, sizeof(float) * qt.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream)); CK(cudaGetLastError());
    for (int a = 0; a < DQ_OUT; ++a) for (int i = 0; i < n; ++i) q[(size_t)i * DQ_OUT + a] = qt[(size_t)a * n + i];     // action-major on the device
    cudaFree(d_pos); cudaFree(d_q);
    return RLPT_OK;
}

static int dqn_hook(void* d_buf, uint64_t count, int dtype, void* stream, void* user) {
    rlpt_ctx* c = (rlpt_ctx*)user;
    return c->allreduce(d_buf, count, dtype, stream, c->allreduce_user);
}
int rlpt_dqn_train_batch(rlpt_ctx* c, const float* pos3, const uint32_t* actions, const float* targets, int n, int apply_update, float* loss) {
    if (!c || !c->dq.ready) return fail(RLPT_ERR_ARG, "rlpt_dqn_train_batch: no network");
    if (n <= 0 || !pos3 || !actions || !targets) return fail(RLPT_ERR_ARG, "rlpt_dqn_train_batch: bad arguments");
    for (int i = 0; i < n; ++i) if (actions[i] >= (uint32_t)DQ_OUT) return fail(RLPT_ERR_ARG, "rlpt_dqn_train_batch: action index out of range");
    CK(cudaSetDevice(c->device));
    std::vector<float4> h_pos(n); for (int i = 0; i < n; ++i) h_pos[i] = make_float4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], 0.f);
    float4* d_pos = nullptr; uint32_t* d_act = nullptr; float* d_tgt = nullptr;
    CK(cudaMalloc(&d_pos, sizeof(float4) * (size_t)n)); CK(cudaMalloc(&d_act, 4 * (size_t)n)); CK(cudaMalloc(&d_tgt, 4 * (size_t)n));
    CK(cudaMemcpyAsync(d_pos, h_pos.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_act, actions, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream)); CK(cudaMemcpyAsync(d_tgt, targets, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    const bool dist = c->allreduce && c->cfg.world_size > 1;
    const int cap0 = c->dq_train.capacity;
    int rc = dqn_train_batch(c->dq, c->dq_train, d_pos, d_act, d_tgt, n, apply_update != 0, dist ? dqn_hook : nullptr, c, c->stream);
    if (c->dq_train.capacity != cap0) nq_graph_reset(c);               // the captured optimiser step points into the buffers that were just reallocated
    float h_loss = 0.f;
    if (!rc) { CK(cudaMemcpyAsync(&h_loss, c->dq_train.scalars, 4, cudaMemcpyDeviceToHost, c->stream)); }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_pos); cudaFree(d_act); cudaFree(d_tgt);
    if (rc) return fail(rc == -2 ? RLPT_ERR_COLLECTIVE : RLPT_ERR_CUDA, rc == -2 ? "gradient all-reduce hook failed" : std::string("rlpt_dqn_train_batch: ") + cudaGetErrorString((cudaError_t)rc));
    if (e != cudaSuccess) return fail(RLPT_ERR_CUDA, std::string("rlpt_dqn_train_batch: ") + cudaGetErrorString(e));
    if (loss) *loss = h_loss;
    return RLPT_OK;
}
// replaces: the training step of NN_Q_Value_Trainer/Source/main.cu:67-135 (fit the network to saved Q tables: squared distance
// over all 144 outputs, summed over the batch, Adam)
int rlpt_dqn_train_supervised(rlpt_ctx* c, const float* pos3, const float* targets144, int n, int apply_update, float* loss) {
    if (!c || !c->dq.ready) return fail(RLPT_ERR_ARG, "rlpt_dqn_train_supervised: no network");
    if (n <= 0 || !pos3 || !targets144) return fail(RLPT_ERR_ARG, "rlpt_dqn_train_supervised: bad arguments");
    CK(cudaSetDevice(c->device));
    std::vector<float4> h_pos(n); for (int i = 0; i < n; ++i) h_pos[i] = make_float4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], 0.f);
    float4* d_pos = nullptr; float* d_tgt = nullptr;
    CK(cudaMalloc(&d_pos, sizeof(float4) * (size_t)n)); CK(cudaMalloc(&d_tgt, 4 * (size_t)n * DQ_OUT));
    CK(cudaMemcpyAsync(d_pos, h_pos.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_tgt, targets144, 4 * (size_t)n * DQ_OUT, cudaMemcpyHostToDevice, c->stream));
    const bool dist = c->allreduce && c->cfg.world_size > 1;
    const int cap0 = c->dq_train.capacity;
    int rc = dqn_train_batch(c->dq, c->dq_train, d_pos, nullptr, d_tgt, n, apply_update != 0, dist ? dqn_hook : nullptr, c, c->stream, true);
    if (c->dq_train.capacity != cap0) nq_graph_reset(c);
    float h_loss = 0.f;
    if (!rc) { CK(cudaMemcpyAsync(&h_loss, c->dq_train.scalars, 4, cudaMemcpyDeviceToHost, c->stream)); }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_pos); cudaFree(d_tgt);
    if (rc) return fail(rc == -2 ? RLPT_ERR_COLLECTIVE : RLPT_ERR_CUDA, rc == -2 ? "gradient all-reduce hook failed" : std::string("rlpt_dqn_train_supervised: ") + cudaGetErrorString((cudaError_t)rc));
    if (e != cudaSuccess) return fail(RLPT_ERR_CUDA, std::string("rlpt_dqn_train_supervised: ") + cudaGetErrorString(e));
    if (loss) *loss = h_loss;
    return RLPT_OK;
}
int rlpt_dqn_get_grads(rlpt_ctx* c, float* grads, int count) {
    if (!c || !c->dq.ready || !c->dq_train.gw[0] || !grads) return fail(RLPT_ERR_ARG, "rlpt_dqn_get_grads: no gradients (call rlpt_dqn_train_batch first)");
    CK(cudaSetDevice(c->device));
    DqnHost shape; shape.k_in = c->dq.k_in;
    if (count != dqn_total(shape)) return fail(RLPT_ERR_ARG, "rlpt_dqn_get_grads: wrong parameter count");
    CK(cudaStreamSynchronize(c->stream));
    float* p = grads;
    for (int l = 0; l < 4; ++l) {
        const size_t nw = (size_t)DqnHost::rows(l) * shape.cols(l), nb = DqnHost::rows(l);
        CK(cudaMemcpy(p, c->dq_train.gw[l], 4 * nw, cudaMemcpyDeviceToHost)); p += nw;
        CK(cudaMemcpy(p, c->dq_train.gb[l], 4 * nb, cudaMemcpyDeviceToHost)); p += nb;
    }
    return RLPT_OK;
}

// ------------------------------------------------------------------------------------------------ frames
// Lanes: the frame's spp are split into L equal slices (L = RLPT_LANES or 2, reduced until it divides spp and every
// slice still holds >= 2^19 paths), each traced on its own stream with its own queues.
static int choose_lanes(const rlpt_config& g) {
    int want = 2;
    if (const char* e = getenv("RLPT_LANES")) want = std::max(1, atoi(e));
    int L = std::min(want, g.spp);
    while (L > 1 && (g.spp % L != 0 || (double)g.width * g.height * (g.spp / L) < 524288.0)) --L;
    return L;
}
static int ensure_frame_buffers(rlpt_ctx* c) {
    const rlpt_config& g = c->cfg;
    const int L = choose_lanes(g), lane_spp = g.spp / L;
    const size_t paths = (size_t)g.width * g.height * lane_spp;
    // NSUB sub-queues of sub_cap slots each (rlpt_kernels.cu "Sub-queues"); counters: one 32-byte sector per (bounce, sub-queue)
    const size_t sub_cap = ((paths + NSUB - 1) / NSUB + 31) / 32 * 32, slots = sub_cap * NSUB;
    const size_t counts_ints = (size_t)(g.max_bounces + 2) * NSUB * COUNT_STRIDE;
    if ((int)c->lanes.size() != L || paths > c->lane_capacity || c->counts_len < g.max_bounces + 2) {
        CK(cudaStreamSynchronize(c->stream));
        for (auto& l : c->lanes) if (l.stream) CK(cudaStreamSynchronize(l.stream));
        free_lanes(c);
        c->lanes.resize(L);
        for (auto& l : c->lanes) {
            int prio_lo = 0, prio_hi = 0; CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CK(cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, prio_hi)); CK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
            CK(cudaStreamCreateWithPriority(&l.pre, cudaStreamNonBlocking, prio_lo));
            CK(cudaEventCreateWithFlags(&l.hit_free, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&l.pre_done, cudaEventDisableTiming));
            for (int k = 0; k < 2; ++k) {
                CK(cudaMalloc(&l.q[k].o, sizeof(float4) * slots)); CK(cudaMalloc(&l.q[k].d, sizeof(float4) * slots));
                CK(cudaMalloc(&l.q[k].thr, sizeof(float4) * slots)); CK(cudaMalloc(&l.q[k].meta, sizeof(uint32_t) * slots));
            }
            CK(cudaMalloc(&l.d_counts, sizeof(int) * counts_ints)); CK(cudaMalloc(&l.d_cursor, sizeof(int) * counts_ints)); CK(cudaMalloc(&l.d_hit, sizeof(float2) * slots));
            CK(cudaHostAlloc(&l.h_counts, sizeof(int) * counts_ints * rlpt_ctx::Lane::SNAPS, cudaHostAllocDefault));
            for (auto& e : l.snap_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        c->lane_capacity = paths; c->counts_len = g.max_bounces + 2;
    }
    c->lane_spp = lane_spp; c->sub_cap = (int)sub_cap;
    if (c->accum_pixels != g.width * g.height) {
        cudaFree(c->d_accum); CK(cudaMalloc(&c->d_accum, sizeof(float4) * (size_t)g.width * g.height));
        CK(cudaMemsetAsync(c->d_accum, 0, sizeof(float4) * (size_t)g.width * g.height, c->stream)); c->accum_pixels = g.width * g.height;
    }
    return RLPT_OK;
}

static cudaEvent_t kev_next(rlpt_ctx* c) {
    if (c->kev_used == c->kev_pool.size()) { cudaEvent_t e = nullptr; if (cudaEventCreate(&e) != cudaSuccess) return nullptr; c->kev_pool.push_back(e); }
    return c->kev_pool[c->kev_used++];
}
static bool kev_room(rlpt_ctx* c) { return c->kev_pending.size() < 16384; }      // beyond that the timings cover a sample of the launches
static cudaEvent_t kev_mark(rlpt_ctx* c, cudaStream_t s) { cudaEvent_t e = kev_next(c); if (e) cudaEventRecord(e, s); return e; }
static void kev_pair(rlpt_ctx* c, cudaEvent_t a, cudaEvent_t b, int kind) { if (a && b) c->kev_pending.push_back({ a, b, kind }); }
// call only when every stream of the context has drained
static void kev_resolve(rlpt_ctx* c) {
    for (const auto& p : c->kev_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { c->k_seconds[p.kind] += ms * 1e-3; c->k_launches[p.kind] += 1.0; }
    }
    (void)cudaGetLastError();
    c->kev_pending.clear(); c->kev_used = 0;
}

// Enqueue one frame of `method` (0 default, 1 SARSA): per lane, primary cast + (max_bounces-1) bounce launches, all
// asynchronous; the live-ray count of every bounce stays on the device. The lanes fork from and join back into the
// context stream with events, so callers still see one ordered stream.
static int enqueue_trace(rlpt_ctx* c, int method, int learn) {
    int rc = ensure_frame_buffers(c); if (rc) return rc;
    const rlpt_config& g = c->cfg;
    FrameDyn dyn{};
    const uint32_t frame_base = (uint32_t)((c->frames_done * (uint64_t)g.world_size + (uint64_t)g.rank) * (uint64_t)g.spp);
    dyn.learn = learn; dyn.cam_x = c->cam[0]; dyn.cam_y = c->cam[1]; dyn.cam_z = c->cam[2];
    dyn.cy = cosf(c->yaw_y); dyn.sy = sinf(c->yaw_y); dyn.cx = cosf(c->yaw_x); dyn.sx = sinf(c->yaw_x);
    dyn.rotated = (c->yaw_y != 0.f || c->yaw_x != 0.f) ? 1 : 0;
    dyn.capture_bounce = c->cap_bounce; dyn.capture_max = c->cap_bounce >= 0 ? c->cap_max : 0; dyn.max_dir = c->max_dir;
    FrameParams p{};
    p.scene = c->scene; p.rm = c->rm; p.accum = c->d_accum; p.stats = c->d_stats;
    p.capture_o = c->d_cap_o; p.capture_d = c->d_cap_d; p.capture_n = c->d_cap_n;
    p.width = g.width; p.height = g.height; p.spp = c->lane_spp; p.max_bounces = g.max_bounces; p.seed = g.seed; p.env = g.env_light;
    const int grid = (c->n_sm * 8 + NSUB - 1) / NSUB * NSUB;              // a whole number of CTAs per sub-queue
    // the split kernels run as one resident wave each (whole CTAs per sub-queue, rounded down)
    if (c->resident_smem != c->smem_bytes || c->per_sm_isect <= 0) { kernels_resident_ctas(c->smem_bytes, c->scene.brute, c->scene.staged, &c->per_sm_isect, &c->per_sm_shade); c->resident_smem = c->smem_bytes; }
    const int per_sm_isect = c->per_sm_isect, per_sm_shade = c->per_sm_shade;
    int grid_isect = std::max(1, c->n_sm * per_sm_isect / NSUB) * NSUB, grid_shade = std::max(1, c->n_sm * per_sm_shade / NSUB) * NSUB;
    if (const char* e = getenv("RLPT_GRID_ISECT")) grid_isect = std::max(1, atoi(e)) * NSUB;
    if (const char* e = getenv("RLPT_GRID_SHADE")) grid_shade = std::max(1, atoi(e)) * NSUB;
    const int split = c->pipe_split, tail_thr = c->pipe_tail;
    const int len = (g.max_bounces + 2) * NSUB * COUNT_STRIDE;
    p.sub_cap = c->sub_cap;
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for (size_t li = 0; li < c->lanes.size(); ++li) {
        rlpt_ctx::Lane& l = c->lanes[li];
        CK(cudaStreamWaitEvent(l.stream, c->ev_fork, 0));
        CK(cudaMemsetAsync(l.d_counts, 0, sizeof(int) * len, l.stream));
        p.q[0] = l.q[0]; p.q[1] = l.q[1]; p.counts = l.d_counts; p.hit = l.d_hit; p.cursor = l.d_cursor;
        dyn.sample_base = frame_base + (uint32_t)li * (uint32_t)c->lane_spp;
        // launch plan: per-bounce launches up to b_tail, then one run-to-completion launch. b_tail = the first bounce the
        // newest arrived snapshot of this lane entered with at most tail_thr live paths (no snapshot yet: every bounce gets
        // its launch). Only speed depends on the choice; the run-to-completion kernel is correct for any count.
        int b_tail = g.max_bounces;
        if (tail_thr > 0 && c->cap_bounce < 0) {
            // stay at most two frames ahead of the device, so that a recent snapshot has always arrived (two frames of
            // queued work keep the GPU busy; an unbounded lead would plan every frame of a long render call blind)
            if (l.snap_seq >= 2) CK(cudaEventSynchronize(l.snap_ev[(l.snap_seq - 2) % rlpt_ctx::Lane::SNAPS]));
            for (uint64_t k = 0; k < rlpt_ctx::Lane::SNAPS && k < l.snap_seq; ++k) {
                const int slot = (int)((l.snap_seq - 1 - k) % rlpt_ctx::Lane::SNAPS);
                if (cudaEventQuery(l.snap_ev[slot]) != cudaSuccess) continue;
                if (l.snap_tag[slot] != g.max_bounces) break;         // taken under another configuration
                const int* hc = l.h_counts + (size_t)slot * len;
                for (int b = 1; b < g.max_bounces; ++b) {
                    long long live = 0; for (int k = 0; k < NSUB; ++k) live += hc[(b * NSUB + k) * COUNT_STRIDE];
                    if (live <= tail_thr) { b_tail = b; break; }
                }
                break;
            }
            (void)cudaGetLastError();
        }
        int launches = 0;
        const bool timed = kev_room(c) && (c->frames_done % 4) == 0;        // every fourth frame: enough for the averages, no event traffic on the rest
        for (int b = 0; b < b_tail; ++b) {
            if (split) {
                cudaStream_t s1 = l.stream;
                if (b == 0 && c->pipe_pre) { s1 = l.pre; if (l.hit_free_valid) CK(cudaStreamWaitEvent(l.pre, l.hit_free, 0)); }
                if (b == 0 && !c->scene.brute) CK(cudaMemsetAsync(l.d_cursor, 0, sizeof(int) * len, s1));       // ray-fetch cursors of k_isect_bvh (free once the previous frame's last k_isect is done)
                cudaEvent_t e0 = timed ? kev_mark(c, s1) : nullptr;
                launch_isect(p, dyn, b, grid_isect, c->smem_bytes, s1);
                cudaEvent_t e1 = timed ? kev_mark(c, s1) : nullptr;
                if (s1 != l.stream) { CK(cudaEventRecord(l.pre_done, l.pre)); CK(cudaStreamWaitEvent(l.stream, l.pre_done, 0)); }
                cudaEvent_t e2 = (timed && s1 != l.stream) ? kev_mark(c, l.stream) : e1;
                launch_shade(p, dyn, method, b, grid_shade, l.stream);
                cudaEvent_t e3 = timed ? kev_mark(c, l.stream) : nullptr;
                kev_pair(c, e0, e1, 0); kev_pair(c, e2, e3, 1); c->k_all[0] += 1.0; c->k_all[1] += 1.0;
                launches += 2;
            } else {
                if (b == 0) launch_primary(p, dyn, method, grid, c->smem_bytes, l.stream); else launch_bounce(p, dyn, method, b, grid, c->smem_bytes, l.stream);
                launches += 1;
            }
        }
        if (split) { CK(cudaEventRecord(l.hit_free, l.stream)); l.hit_free_valid = true; }
        if (b_tail < g.max_bounces) {
            cudaEvent_t e0 = timed ? kev_mark(c, l.stream) : nullptr;
            launch_tail(p, dyn, method, b_tail, grid, c->smem_bytes, l.stream); launches += 1;
            cudaEvent_t e1 = timed ? kev_mark(c, l.stream) : nullptr;
            kev_pair(c, e0, e1, 2); c->k_all[2] += 1.0;
        }
        c->launches += (double)launches;
        {   // snapshot of this frame's counts
            const int slot = (int)(l.snap_seq % rlpt_ctx::Lane::SNAPS);
            CK(cudaMemcpyAsync(l.h_counts + (size_t)slot * len, l.d_counts, sizeof(int) * len, cudaMemcpyDeviceToHost, l.stream));
            CK(cudaEventRecord(l.snap_ev[slot], l.stream)); l.snap_tag[slot] = g.max_bounces; l.snap_seq++;
        }
        CK(cudaEventRecord(l.done, l.stream));
        CK(cudaStreamWaitEvent(c->stream, l.done, 0));
    }
    CK(cudaGetLastError());
    c->frames_done++;
    c->cap_bounce = -1;
    return RLPT_OK;
}

static int enqueue_merge(rlpt_ctx* c) {
    if (c->p2p_ready && c->cfg.world_size > 1) {
        launch_merge_p2p(c->rm, c->peers, c->d_surf_lum_over_pi, c->cfg.radiance_threshold, ++c->p2p_epoch, c->d_p2p_done, c->stream);
        c->launches += 2; c->p2p_used = true;
        CK(cudaGetLastError());
        return RLPT_OK;
    }
    if (c->allreduce && c->cfg.world_size > 1) {
        const uint64_t cells = (uint64_t)c->rm.n_vol * CELLS;
        if (c->allreduce(c->d_acc_sum, cells, 0, (void*)c->stream, c->allreduce_user)) return fail(RLPT_ERR_COLLECTIVE, "all-reduce hook failed (target sums)");
        if (c->allreduce(c->d_acc_cnt, cells, 1, (void*)c->stream, c->allreduce_user)) return fail(RLPT_ERR_COLLECTIVE, "all-reduce hook failed (visit counts)");
    }
    launch_merge(c->rm, c->d_surf_lum_over_pi, c->cfg.radiance_threshold, 0, c->stream);
    c->launches += 1;
    CK(cudaGetLastError());
    return RLPT_OK;
}

// phase marks: events recorded in stream order around the tracing kernels and around the merge of every frame of one
// render call; read back (after the call's final synchronisation) into trace_seconds / merge_seconds
static int phase_mark(rlpt_ctx* c) {
    if (c->phase_used == c->phase_ev.size()) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->phase_ev.push_back(e); }
    CK(cudaEventRecord(c->phase_ev[c->phase_used++], c->stream));
    return RLPT_OK;
}
static int timed_begin(rlpt_ctx* c) { CK(cudaSetDevice(c->device)); c->phase_used = 0; CK(cudaEventRecord(c->ev0, c->stream)); return RLPT_OK; }
static int timed_end(rlpt_ctx* c, int frames) {
    CK(cudaEventRecord(c->ev1, c->stream)); CK(cudaEventSynchronize(c->ev1));
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->device_seconds += ms * 1e-3; c->frames_rendered += frames;
    // marks come in (trace begin, trace end[, merge end]) groups of `marks_per_frame`
    const size_t per = frames > 0 ? c->phase_used / (size_t)frames : 0;
    for (size_t f = 0; per >= 2 && f < (size_t)frames; ++f) {
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, c->phase_ev[f * per], c->phase_ev[f * per + 1])); c->trace_seconds += a * 1e-3;
        if (per >= 3) { CK(cudaEventElapsedTime(&b, c->phase_ev[f * per + 1], c->phase_ev[f * per + 2])); c->merge_seconds += b * 1e-3; }
    }
    c->phase_used = 0;
    kev_resolve(c);
    CK(cudaGetLastError());
    return p2p_check(c);
}

int rlpt_render_default(rlpt_ctx* c, int frames) {
    if (!c || !c->have_scene) return fail(RLPT_ERR_ARG, "rlpt_render_default: upload a scene first");
    if (frames < 0) return fail(RLPT_ERR_ARG, "rlpt_render_default: negative frame count");
    int rc = timed_begin(c); if (rc) return rc;
    for (int f = 0; f < frames; ++f) { rc = phase_mark(c); if (rc) return rc; rc = enqueue_trace(c, 0, 0); if (rc) return rc; rc = phase_mark(c); if (rc) return rc; }
    return timed_end(c, frames);
}
int rlpt_sarsa_trace(rlpt_ctx* c) {
    if (!c || !c->have_scene || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_sarsa_trace: needs a scene and a radiance map");
    CK(cudaSetDevice(c->device));
    return enqueue_trace(c, 1, 1);
}
int rlpt_sarsa_merge(rlpt_ctx* c) {
    if (!c || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_sarsa_merge: no radiance map");
    CK(cudaSetDevice(c->device));
    return enqueue_merge(c);
}
int rlpt_render_sarsa(rlpt_ctx* c, int frames) {
    if (!c || !c->have_scene || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_render_sarsa: needs a scene and a radiance map");
    if (frames < 0) return fail(RLPT_ERR_ARG, "rlpt_render_sarsa: negative frame count");
    int rc = timed_begin(c); if (rc) return rc;
    for (int f = 0; f < frames; ++f) {
        rc = phase_mark(c); if (rc) return rc;
        rc = enqueue_trace(c, 1, 1); if (rc) return rc;
        rc = phase_mark(c); if (rc) return rc;
        rc = enqueue_merge(c); if (rc) return rc;
        rc = phase_mark(c); if (rc) return rc;
    }
    return timed_end(c, frames);
}
// replaces: the VORONOI method of main (G/main.cu:413-470) -> draw_voronoi_trace (G/path_tracing/voronoi_trace.cu:4-45)
int rlpt_render_voronoi(rlpt_ctx* c) {
    if (!c || !c->have_scene || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_render_voronoi: needs a scene and a radiance map");
    CK(cudaSetDevice(c->device));
    int rc = ensure_frame_buffers(c); if (rc) return rc;
    const rlpt_config& g = c->cfg;
    FrameDyn dyn{};
    dyn.sample_base = (uint32_t)c->frames_done; dyn.cam_x = c->cam[0]; dyn.cam_y = c->cam[1]; dyn.cam_z = c->cam[2];
    dyn.cy = cosf(c->yaw_y); dyn.sy = sinf(c->yaw_y); dyn.cx = cosf(c->yaw_x); dyn.sx = sinf(c->yaw_x);
    dyn.rotated = (c->yaw_y != 0.f || c->yaw_x != 0.f) ? 1 : 0;
    FrameParams p{};
    p.scene = c->scene; p.rm = c->rm; p.accum = c->d_accum; p.stats = c->d_stats;
    p.width = g.width; p.height = g.height; p.spp = 1; p.max_bounces = g.max_bounces; p.seed = g.seed; p.env = g.env_light;
    launch_voronoi(p, dyn, c->n_sm * 8, c->smem_bytes, c->stream);
    c->launches += 1; c->frames_done++;
    CK(cudaGetLastError());
    return RLPT_OK;
}
int rlpt_render_sarsa_frozen(rlpt_ctx* c, int frames) {
    if (!c || !c->have_scene || !c->have_rmap) return fail(RLPT_ERR_ARG, "rlpt_render_sarsa_frozen: needs a scene and a radiance map");
    if (frames < 0) return fail(RLPT_ERR_ARG, "negative frame count");
    int rc = timed_begin(c); if (rc) return rc;
    for (int f = 0; f < frames; ++f) { rc = phase_mark(c); if (rc) return rc; rc = enqueue_trace(c, 1, 0); if (rc) return rc; rc = phase_mark(c); if (rc) return rc; }
    return timed_end(c, frames);
}

// replaces: PretrainedPathtracer (G/deep_learning/pre_trained_pathtracer.cu:10-185 host loop, :186-378 render_frame). The
// frame's samples are traced in slices of at most 2^21 paths (bounds the Q buffer: 144 floats per live path).
static int enqueue_nq_inference(rlpt_ctx* c) {
    int rc = ensure_frame_buffers(c); if (rc) return rc;
    const rlpt_config& g = c->cfg;
    int slice = c->lane_spp;
    while (slice > 1 && ((double)g.width * g.height * slice > 2097152.0 || c->lane_spp % slice != 0)) --slice;
    const size_t cap = (size_t)g.width * g.height * slice;
    if (cap > c->nq_q_capacity) { cudaFree(c->d_nq_q); c->d_nq_q = nullptr; c->nq_q_capacity = 0; CK(cudaMalloc(&c->d_nq_q, sizeof(float) * DQ_OUT * cap)); c->nq_q_capacity = cap; }
    FrameDyn dyn{};
    const uint32_t frame_base = (uint32_t)((c->frames_done * (uint64_t)g.world_size + (uint64_t)g.rank) * (uint64_t)g.spp);
    dyn.cam_x = c->cam[0]; dyn.cam_y = c->cam[1]; dyn.cam_z = c->cam[2];
    dyn.cy = cosf(c->yaw_y); dyn.sy = sinf(c->yaw_y); dyn.cx = cosf(c->yaw_x); dyn.sx = sinf(c->yaw_x);
    dyn.rotated = (c->yaw_y != 0.f || c->yaw_x != 0.f) ? 1 : 0;
    dyn.max_dir = c->max_dir;
    rlpt_ctx::Lane& l = c->lanes[0];
    FrameParams p{};
    p.scene = c->scene; p.rm = c->rm; p.accum = c->d_accum; p.stats = c->d_stats; p.q[0] = l.q[0]; p.q[1] = l.q[1]; p.counts = l.d_counts;
    p.width = g.width; p.height = g.height; p.spp = slice; p.max_bounces = g.max_bounces; p.seed = g.seed; p.env = g.env_light;
    DqnFwdParams fp{}; fp.c1 = c->dq.c1; fp.m1 = c->dq.m1; fp.b2 = c->dq.b[1]; fp.b3 = c->dq.b[2]; fp.b4 = c->dq.b[3];
    fp.w2p = c->dq.w2p; fp.w3p = c->dq.w3p; fp.w4p = c->dq.w4p; fp.q = c->d_nq_q; fp.q_stride = (int)cap; fp.n = (int)cap;
    const int grid = c->n_sm * 8;
    for (int s0 = 0; s0 < g.spp; s0 += slice) {
        dyn.sample_base = frame_base + (uint32_t)s0;
        CK(cudaMemsetAsync(l.d_counts, 0, sizeof(int) * (g.max_bounces + 2), c->stream));
        launch_nq_trace(p, dyn, 0, grid, c->smem_bytes, c->stream);
        for (int b = 1; b < g.max_bounces; ++b) {
            fp.pos = p.q[b & 1].o; fp.n_ptr = l.d_counts + b;
            int frc = dqn_forward(c->dq, fp, c->stream);
            if (frc) return fail(RLPT_ERR_CUDA, std::string("DQN forward launch failed: ") + cudaGetErrorString((cudaError_t)frc));
            launch_nq_sample(p, dyn, b, c->d_nq_q, (int)cap, 0.f, nullptr, grid, c->stream);
            launch_nq_trace(p, dyn, b, grid, c->smem_bytes, c->stream);
        }
        c->launches += 1.0 + 3.0 * (g.max_bounces - 1);
    }
    CK(cudaGetLastError());
    c->frames_done++;
    return RLPT_OK;
}
int rlpt_render_pretrained(rlpt_ctx* c, int frames) {
    if (!c || !c->have_scene || !c->dq.ready) return fail(RLPT_ERR_ARG, "rlpt_render_pretrained: needs a scene and a network (rlpt_dqn_load_text / rlpt_dqn_init)");
    if (frames < 0) return fail(RLPT_ERR_ARG, "rlpt_render_pretrained: negative frame count");
    int rc = timed_begin(c); if (rc) return rc;
    for (int f = 0; f < frames; ++f) { rc = phase_mark(c); if (rc) return rc; rc = enqueue_nq_inference(c); if (rc) return rc; rc = phase_mark(c); if (rc) return rc; }
    return timed_end(c, frames);
}

// replaces: NeuralQPathtracer (G/deep_learning/neural_q_pathtracer.cu:5-223 host loop, :226-600 render_frame). One frame =
// cfg.spp passes over all pixels; in every pass each bounce samples directions from the network (epsilon-greedy), traces all
// rays, and trains the network on the transitions in batches of `batch` rays (sequential optimiser steps, as the reference).
// The live-ray count is read back once per bounce (the reference's rays_finished flag, :536).
static int ensure_nqt(rlpt_ctx* c, int batch) {
    const rlpt_config& g = c->cfg; const int n = g.width * g.height;
    if (c->nqt_n == n && c->nqt_batch == batch && c->nqt.loc) return RLPT_OK;
    CK(cudaStreamSynchronize(c->stream)); free_nqt(c);
    NqTrainState& st = c->nqt; st.n = n;
    CK(cudaMalloc(&st.loc, sizeof(float4) * n)); CK(cudaMalloc(&st.sloc, sizeof(float4) * n)); CK(cudaMalloc(&st.dir, sizeof(float4) * n)); CK(cudaMalloc(&st.thr, sizeof(float4) * n));
    CK(cudaMalloc(&st.state, 4 * (size_t)n)); CK(cudaMalloc(&st.reward, 4 * (size_t)n)); CK(cudaMalloc(&st.discount, 4 * (size_t)n)); CK(cudaMalloc(&st.action, 4 * (size_t)n));
    CK(cudaMalloc(&st.alive, 4 * (size_t)(255 + 2)));               // rlpt_config_set accepts max_bounces up to 255; the buffers are reused across such changes
    const int S = (batch + DQ_TILE - 1) / DQ_TILE * DQ_TILE;
    CK(cudaMalloc(&c->d_nqt_qcur, sizeof(float) * DQ_OUT * (size_t)n)); CK(cudaMalloc(&c->d_nqt_qnext, sizeof(float) * DQ_OUT * (size_t)S));
    CK(cudaMalloc(&c->d_nqt_targets, 4 * (size_t)S)); CK(cudaMalloc(&c->d_nqt_loss, 4)); CK(cudaMemset(c->d_nqt_loss, 0, 4));
    c->nqt_n = n; c->nqt_batch = batch;
    return RLPT_OK;
}
static int enqueue_nq_training_frame(rlpt_ctx* c, int batch) {
    const rlpt_config& g = c->cfg; const int n = g.width * g.height;
    c->dq_train.lr = c->nq_lr;                                 // (the optimiser state is recreated whenever new parameters are pushed)
    int rc = ensure_nqt(c, batch); if (rc) return rc;
    if (c->accum_pixels != n) {
        cudaFree(c->d_accum); CK(cudaMalloc(&c->d_accum, sizeof(float4) * (size_t)n));
        CK(cudaMemsetAsync(c->d_accum, 0, sizeof(float4) * (size_t)n, c->stream)); c->accum_pixels = n;
    }
    FrameDyn dyn{};
    dyn.cam_x = c->cam[0]; dyn.cam_y = c->cam[1]; dyn.cam_z = c->cam[2];
    dyn.cy = cosf(c->yaw_y); dyn.sy = sinf(c->yaw_y); dyn.cx = cosf(c->yaw_x); dyn.sx = sinf(c->yaw_x);
    dyn.rotated = (c->yaw_y != 0.f || c->yaw_x != 0.f) ? 1 : 0;
    FrameParams p{};
    p.scene = c->scene; p.accum = c->d_accum; p.stats = c->d_stats;
    p.width = g.width; p.height = g.height; p.spp = 1; p.max_bounces = g.max_bounces; p.seed = g.seed; p.env = g.env_light;
    DqnFwdParams fp{}; fp.c1 = c->dq.c1; fp.m1 = c->dq.m1; fp.b2 = c->dq.b[1]; fp.b3 = c->dq.b[2]; fp.b4 = c->dq.b[3]; fp.w2p = c->dq.w2p; fp.w3p = c->dq.w3p; fp.w4p = c->dq.w4p;
    const int grid = c->n_sm * 8;
    const bool dist = c->allreduce && g.world_size > 1;
    const uint32_t frame_base = (uint32_t)((c->frames_done * (uint64_t)g.world_size + (uint64_t)g.rank) * (uint64_t)g.spp);
    const int S = (batch + DQ_TILE - 1) / DQ_TILE * DQ_TILE;
    for (int pass = 0; pass < g.spp; ++pass) {
        dyn.sample_base = frame_base + (uint32_t)pass;
        CK(cudaMemsetAsync(c->nqt.alive, 0, 4 * (size_t)(g.max_bounces + 2), c->stream));
        launch_nqt_init(p, dyn, c->nqt, c->stream);
        for (int b = 0; b < g.max_bounces; ++b) {
            if (b > 0) {
                fp.pos = c->nqt.loc; fp.n = n; fp.n_ptr = nullptr; fp.q = c->d_nqt_qcur; fp.q_stride = n; fp.h1t = fp.h2t = fp.h3t = nullptr;
                cudaEvent_t e0 = kev_room(c) ? kev_mark(c, c->stream) : nullptr;
                int frc = dqn_forward(c->dq, fp, c->stream); if (frc) return fail(RLPT_ERR_CUDA, "DQN forward launch failed");
                cudaEvent_t e1 = e0 ? kev_mark(c, c->stream) : nullptr;
                kev_pair(c, e0, e1, 3); c->k_all[3] += 1.0; c->dqn_rays += (double)n;
                launch_nqt_sample(p, dyn, c->nqt, b, c->d_nqt_qcur, n, c->nq_epsilon, c->stream);
            }
            launch_nqt_trace(p, dyn, c->nqt, b, grid, c->smem_bytes, c->stream);
            c->launches += b > 0 ? 3.0 : 1.0;
            if (b > 0) {
                cudaEvent_t t0 = kev_room(c) ? kev_mark(c, c->stream) : nullptr;
                // One optimiser step is ten small launches (step begin, forward of states and next states, backward data path, three weight-gradient
                // GEMMs, dW4 scatter, gather + norm, Adam + operands): launch-bound. All FULL batches of a bounce replay ONE CUDA graph captured once
                // (the batch offsets are the same every bounce, so the nodes read the ray arrays in place -- no staging copy, one graph launch per
                // bounce instead of one per step); a ragged last batch, and multi-GPU runs (the all-reduce hook is a host callback), go step by step.
                const int full = (c->nq_graphs && !dist) ? n / batch : 0;
                auto one_step = [&](int start, int bn) -> int {
                    int frc = dqn_train_begin(c->dq, c->dq_train, c->nqt.sloc + start, bn, true, c->stream); if (frc) return frc;
                    // the step's zeroing runs beside the forward launch, which evaluates the batch's states (activations kept for the backward pass) and
                    // its next states (Q only); the TD targets are derived from those inside the backward kernel
                    DqnFwdParams gp = dqn_train_forward_params(c->dq, c->dq_train, c->nqt.sloc + start, bn);
                    gp.pos2 = c->nqt.loc + start; gp.n2 = bn; gp.q2 = c->d_nqt_qnext; gp.q_stride2 = S;
                    frc = dqn_forward(c->dq, gp, c->stream); if (frc) return frc;
                    const DqnTdParams td{ c->d_nqt_qnext, S, c->nqt.state + start, c->nqt.reward + start, c->nqt.discount + start };
                    return dqn_train_batch(c->dq, c->dq_train, c->nqt.sloc + start, c->nqt.action + start, c->d_nqt_targets, bn, true, dist ? dqn_hook : nullptr, c, c->stream, false, c->d_nqt_loss, true, &td);
                };
                if (full > 0) {
                    if (!c->nq_graph_exec || c->nq_graph_batch != batch) {
                        if (c->nq_graph_exec) { cudaGraphExecDestroy(c->nq_graph_exec); c->nq_graph_exec = nullptr; }
                        int arc = dqn_train_prepare(c->dq, c->dq_train, batch, c->stream); if (arc) return fail(RLPT_ERR_CUDA, "Neural-Q training buffers");
                        cudaGraph_t graph = nullptr; int frc = 0;
                        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                        for (int k = 0; k < full && !frc; ++k) frc = one_step(k * batch, batch);
                        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
                        if (frc || ce != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); c->dq_train.begun = false; return fail(RLPT_ERR_CUDA, "Neural-Q training steps: graph capture failed"); }
                        ce = cudaGraphInstantiate(&c->nq_graph_exec, graph, 0);
                        cudaGraphDestroy(graph);
                        if (ce != cudaSuccess) { c->nq_graph_exec = nullptr; return fail(RLPT_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
                        c->nq_graph_batch = batch;
                    }
                    CK(cudaGraphLaunch(c->nq_graph_exec, c->stream));
                    c->dq_train.step += (unsigned long long)full;
                    c->launches += (c->dq_train.fused_bwd ? 9.0 : 13.0) * full; c->k_all[4] += (double)full;
                }
                for (int start = full * batch; start < n; start += batch) {
                    const int bn = std::min(batch, n - start);
                    int frc = one_step(start, bn);
                    if (frc) return fail(frc == -2 ? RLPT_ERR_COLLECTIVE : RLPT_ERR_CUDA, "Neural-Q training step failed");
                    c->launches += c->dq_train.fused_bwd ? 9.0 : 13.0;       // kernels of one optimiser step: step begin, forward of both batches, backward data path (one kernel, or delta + 2 GEMMs + 2 masks), 3 weight-gradient GEMMs, dW4 scatter, gather + norm, Adam + operand refresh
                    c->k_all[4] += 1.0;
                }
                cudaEvent_t t1 = t0 ? kev_mark(c, c->stream) : nullptr;
                kev_pair(c, t0, t1, 4);
            }
            launch_nqt_respawn(p, dyn, c->nqt, b, c->stream);
            int alive = 0;
            CK(cudaMemcpyAsync(&alive, c->nqt.alive + b + 1, 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if (alive == 0) break;
        }
        c->nq_epsilon = std::max(c->nq_epsilon - c->nq_eps_decay, c->nq_eps_min);               // EPSILON_DECAY, EPSILON_MIN (deep_learning_settings.h:6-7)
    }
    float loss = 0.f; CK(cudaMemcpy(&loss, c->d_nqt_loss, 4, cudaMemcpyDeviceToHost)); CK(cudaMemset(c->d_nqt_loss, 0, 4));
    c->nq_loss_total = loss;
    CK(cudaGetLastError());
    c->frames_done++;
    return RLPT_OK;
}
int rlpt_render_neuralq(rlpt_ctx* c, int frames, int batch) {
    if (!c || !c->have_scene || !c->dq.ready) return fail(RLPT_ERR_ARG, "rlpt_render_neuralq: needs a scene and a network (rlpt_dqn_init / rlpt_dqn_load_text)");
    if (frames < 0 || batch <= 0) return fail(RLPT_ERR_ARG, "rlpt_render_neuralq: bad frame count / batch size");
    int rc = timed_begin(c); if (rc) return rc;
    for (int f = 0; f < frames; ++f) { rc = phase_mark(c); if (rc) return rc; rc = enqueue_nq_training_frame(c, batch); if (rc) return rc; rc = phase_mark(c); if (rc) return rc; }
    return timed_end(c, frames);
}
int rlpt_neuralq_set_hyper(rlpt_ctx* c, float lr, float eps_start, float eps_decay, float eps_min) {
    if (!c) return fail(RLPT_ERR_ARG, "null ctx");
    if (!(lr >= 0.f) || !(eps_start >= 0.f && eps_start <= 1.f) || !(eps_decay >= 0.f) || !(eps_min >= 0.f && eps_min <= 1.f)) return fail(RLPT_ERR_ARG, "rlpt_neuralq_set_hyper: learning rate >= 0, epsilons in [0, 1]");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    nq_graph_reset(c);                                     // the captured optimiser step carries the learning rate by value
    c->nq_lr = lr; c->dq_train.lr = lr; c->nq_epsilon = eps_start; c->nq_eps_decay = eps_decay; c->nq_eps_min = eps_min;
    return RLPT_OK;
}
int rlpt_neuralq_last_loss(rlpt_ctx* c, double* loss) { if (!c || !loss) return fail(RLPT_ERR_ARG, "null"); *loss = c->nq_loss_total; return RLPT_OK; }

int rlpt_frame_reset(rlpt_ctx* c) {
    if (!c) return fail(RLPT_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    if (c->d_accum) CK(cudaMemsetAsync(c->d_accum, 0, sizeof(float4) * (size_t)c->accum_pixels, c->stream));
    return RLPT_OK;
}
int rlpt_frame_allreduce(rlpt_ctx* c) {
    if (!c || !c->d_accum) return fail(RLPT_ERR_ARG, "rlpt_frame_allreduce: nothing rendered");
    if (!c->allreduce || c->cfg.world_size <= 1) return RLPT_OK;
    CK(cudaSetDevice(c->device));
    if (c->allreduce(c->d_accum, (uint64_t)c->accum_pixels * 4, 0, (void*)c->stream, c->allreduce_user)) return fail(RLPT_ERR_COLLECTIVE, "all-reduce hook failed (frame buffer)");
    return RLPT_OK;
}
int rlpt_frame_download(rlpt_ctx* c, float* rgb) {
    if (!c || !c->d_accum || !rgb) return fail(RLPT_ERR_ARG, "rlpt_frame_download: nothing rendered / null");
    CK(cudaSetDevice(c->device));
    size_t n = (size_t)c->accum_pixels;
    int rc = ensure_stage(c, sizeof(float) * 3 * n); if (rc) return rc;
    float* d = (float*)c->d_stage;
    launch_frame_mean(c->d_accum, d, (int)n, c->stream);
    CK(cudaMemcpyAsync(rgb, d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream)); CK(cudaGetLastError());
    return RLPT_OK;
}
int rlpt_frame_download_argb(rlpt_ctx* c, uint32_t* argb) {
    if (!c || !c->d_accum || !argb) return fail(RLPT_ERR_ARG, "rlpt_frame_download_argb: nothing rendered / null");
    CK(cudaSetDevice(c->device));
    size_t n = (size_t)c->accum_pixels;
    int rc = ensure_stage(c, sizeof(uint32_t) * n); if (rc) return rc;
    uint32_t* d = (uint32_t*)c->d_stage;
    launch_pack_argb(c->d_accum, d, c->cfg.width, c->cfg.height, c->stream);
    CK(cudaMemcpyAsync(argb, d, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream)); CK(cudaGetLastError());
    return RLPT_OK;
}

// 32-bpp BMP exactly as SDL_SaveBMP writes the reference's ARGB8888 surface (Images/render.bmp): BITMAPV4HEADER (108 bytes),
// BI_BITFIELDS, masks R 00ff0000 G 0000ff00 B 000000ff A ff000000, pixel data at offset 122, rows bottom-up.
int rlpt_frame_save_bmp(rlpt_ctx* c, const char* path) {
    if (!c || !path) return fail(RLPT_ERR_ARG, "rlpt_frame_save_bmp: null");
    const int w = c->cfg.width, h = c->cfg.height;
    std::vector<uint32_t> px((size_t)w * h);
    int rc = rlpt_frame_download_argb(c, px.data()); if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(RLPT_ERR_IO, std::string("rlpt_frame_save_bmp: cannot open ") + path);
    auto u16 = [&](uint16_t v) { fwrite(&v, 2, 1, f); }; auto u32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
    const uint32_t data = (uint32_t)w * h * 4, off = 14 + 108;
    fputc('B', f); fputc('M', f); u32(off + data); u16(0); u16(0); u32(off);
    u32(108); u32((uint32_t)w); u32((uint32_t)h); u16(1); u16(32); u32(3 /*BI_BITFIELDS*/); u32(data); u32(0); u32(0); u32(0); u32(0);
    u32(0x00ff0000u); u32(0x0000ff00u); u32(0x000000ffu); u32(0xff000000u);
    u32(0x57696e20u /*LCS_WINDOWS_COLOR_SPACE*/); for (int i = 0; i < 9; ++i) u32(0); u32(0); u32(0); u32(0);
    for (int y = h - 1; y >= 0; --y) fwrite(&px[(size_t)y * w], 4, (size_t)w, f);
    fclose(f);
    return RLPT_OK;
}

int rlpt_stats(rlpt_ctx* c, rlpt_stats_t* out) {
    if (!c || !out) return fail(RLPT_ERR_ARG, "rlpt_stats: null");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    kev_resolve(c);
    unsigned long long h[8]; CK(cudaMemcpy(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    out->path_length_sum = (double)h[0]; out->zero_contribution_paths = (double)h[1]; out->paths = (double)h[2];
    out->ray_casts = (double)h[0]; out->device_seconds = c->device_seconds; out->frames = c->frames_rendered; out->kernel_launches = c->launches;
    out->triangle_tests = (double)h[3]; out->box_tests = (double)h[4]; out->trace_seconds = c->trace_seconds; out->merge_seconds = c->merge_seconds;
    out->kd_fallbacks = (double)h[5];
    // event pairs are recorded on every fourth frame; the sums are scaled to all launches of the period
    auto scaled = [&](int k) { return c->k_launches[k] > 0.0 ? c->k_seconds[k] * (c->k_all[k] / c->k_launches[k]) : 0.0; };
    out->isect_seconds = scaled(0); out->isect_launches = c->k_all[0]; out->shade_seconds = scaled(1); out->shade_launches = c->k_all[1];
    out->tail_seconds = scaled(2); out->tail_launches = c->k_all[2];
    out->dqn_forward_seconds = scaled(3); out->dqn_forward_launches = c->k_all[3]; out->dqn_forward_rays = c->dqn_rays;
    out->train_seconds = c->k_seconds[4]; out->train_steps = c->k_all[4];
    return RLPT_OK;
}
int rlpt_stats_reset(rlpt_ctx* c) {
    if (!c) return fail(RLPT_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->d_stats, 0, sizeof(unsigned long long) * 8, c->stream));
    c->device_seconds = 0.0; c->frames_rendered = 0.0; c->launches = 0.0; c->trace_seconds = 0.0; c->merge_seconds = 0.0;
    CK(cudaStreamSynchronize(c->stream)); kev_resolve(c);
    for (int k = 0; k < 5; ++k) { c->k_seconds[k] = 0.0; c->k_launches[k] = 0.0; c->k_all[k] = 0.0; }
    c->dqn_rays = 0.0;
    return RLPT_OK;
}

int rlpt_measure_fp32_peak(rlpt_ctx* c, double* tflops) {
    if (!c || !tflops) return fail(RLPT_ERR_ARG, "null");
    CK(cudaSetDevice(c->device));
    const int grid = c->n_sm * 8, iters = 4096; float* d = nullptr;
    CK(cudaMalloc(&d, sizeof(float) * (size_t)grid * 256));
    launch_fp32_peak(d, 64, grid, c->stream);                         // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(c->ev0, c->stream));
        launch_fp32_peak(d, iters, grid, c->stream);
        CK(cudaEventRecord(c->ev1, c->stream)); CK(cudaEventSynchronize(c->ev1));
        float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        double flops = (double)grid * 256.0 * iters * 16.0 * 8.0 * 2.0;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaFree(d);
    *tflops = best;
    return RLPT_OK;
}

int rlpt_capture_rays(rlpt_ctx* c, int method, int bounce, float* org, float* dir, int max_rays, int* n_captured) {
    if (!c || !c->have_scene || (method == 1 && !c->have_rmap)) return fail(RLPT_ERR_ARG, "rlpt_capture_rays: needs a scene (and a radiance map for method 1)");
    if (max_rays <= 0 || bounce < 0 || !org || !dir || !n_captured) return fail(RLPT_ERR_ARG, "rlpt_capture_rays: bad arguments");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_cap_o); cudaFree(c->d_cap_d); c->d_cap_o = c->d_cap_d = nullptr;
    CK(cudaMalloc(&c->d_cap_o, sizeof(float4) * (size_t)max_rays)); CK(cudaMalloc(&c->d_cap_d, sizeof(float4) * (size_t)max_rays));
    CK(cudaMemset(c->d_cap_n, 0, sizeof(int)));
    c->cap_max = max_rays; c->cap_bounce = bounce;
    // trace one frame that is not counted: no learning, frame counter and accumulators restored afterwards
    uint64_t saved_frames = c->frames_done;
    std::vector<float4> keep; if (c->d_accum) { keep.resize((size_t)c->accum_pixels); CK(cudaMemcpy(keep.data(), c->d_accum, sizeof(float4) * keep.size(), cudaMemcpyDeviceToHost)); }
    unsigned long long hs[8]; CK(cudaMemcpy(hs, c->d_stats, sizeof hs, cudaMemcpyDeviceToHost));
    int rc = enqueue_trace(c, method, 0); if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream)); CK(cudaGetLastError());
    c->frames_done = saved_frames; c->launches -= (double)c->cfg.max_bounces * (double)c->lanes.size();
    if (!keep.empty()) CK(cudaMemcpy(c->d_accum, keep.data(), sizeof(float4) * keep.size(), cudaMemcpyHostToDevice));
    else CK(cudaMemset(c->d_accum, 0, sizeof(float4) * (size_t)c->accum_pixels));
    CK(cudaMemcpy(c->d_stats, hs, sizeof hs, cudaMemcpyHostToDevice));
    int n = 0; CK(cudaMemcpy(&n, c->d_cap_n, sizeof(int), cudaMemcpyDeviceToHost)); n = std::min(n, max_rays);
    std::vector<float4> o(n), d(n);
    if (n) { CK(cudaMemcpy(o.data(), c->d_cap_o, sizeof(float4) * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(d.data(), c->d_cap_d, sizeof(float4) * n, cudaMemcpyDeviceToHost)); }
    for (int i = 0; i < n; ++i) { org[3 * i] = o[i].x; org[3 * i + 1] = o[i].y; org[3 * i + 2] = o[i].z; dir[3 * i] = d[i].x; dir[3 * i + 1] = d[i].y; dir[3 * i + 2] = d[i].z; }
    *n_captured = n;
    return RLPT_OK;
}

}  // extern "C"
