/* oracle/ref_harness.cu -- TEST INFRASTRUCTURE ONLY (checker, never the product, never shipped).
 *
 * A small extern "C" driver around the UNMODIFIED reference engine (GPU_Rendering_Engine/Source/, compiled
 * from where it lies under /root/reference by oracle/build_ref.sh; nothing from it is copied here). It stands
 * in for the reference's own main.cu, which cannot be built (DyNet headers main.cu:28-30, SDL window :94).
 *
 * Two builds of this one file, both written to oracle/_ref/ (git-ignored, but it travels to the GPU box):
 *   libref_cuda.so  nvcc -rdc=true, sm_100a: the reference's kernels exactly as the reference builds them
 *                   (nvcc defaults: -fmad=true, no fast-math; flags.make `CUDA_FLAGS =`). GPU box only.
 *   libref_host.so  g++ with oracle/host_shim: the same sources as host C++, launch grids walked with OpenMP.
 *                   Runs anywhere; pins oracle/rlpt_oracle.cpp and is bench.py's `--impl reference` arm.
 *
 * What each entry point drives (reference file:line):
 *   ref_scene_cornell      Scene::load_cornell_box_scene                      scenes/scene.cu:8-30
 *   ref_scene_obj          Scene::load_custom_scene                           scenes/scene.cu:33-60
 *   ref_closest_hit        Ray::Ray + Ray::closest_intersection               rays/ray.cu:6-36
 *   ref_rmap_build         RadianceMap::RadianceMap (rand() sampling, kd-tree) radiance_volumes/radiance_map.cu:8-54
 *   ref_find_closest       RadianceMap::find_closest_radiance_volume_iterative radiance_map.cu:150-203
 *   ref_rmap_update_distributions  update_radiance_volume_distributions       path_tracing/reinforcement_path_tracing.cu:6-13
 *   ref_grid_dir           convert_grid_pos_to_direction                      utils/hemisphere_helpers.cu:96-105
 *   ref_render_default     main.cu:190-244 frame loop (method 0)
 *   ref_render_sarsa       main.cu:246-364 frame loop (method 1)
 * The frame loops mirror main.cu's call order; cross-frame averaging (which the reference never does,
 * main.cu:359) happens here on the host with a NaN guard, see SURVEY.md section 7.
 */
#include <vector>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <chrono>

#include "scene.cuh"
#include "camera.cuh"
#include "ray.cuh"
#include "radiance_map.cuh"
#include "radiance_tree.cuh"
#include "radiance_volume.cuh"
#include "default_path_tracing.cuh"
#include "reinforcement_path_tracing.cuh"
#include "hemisphere_helpers.cuh"
#include "image_settings.h"
#include "monte_carlo_settings.h"
#include "radiance_volumes_settings.h"

#ifdef RLPT_REF_HOST
#include <omp.h>
thread_local rlpt_uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
template <class F> static void rlpt_walk_grid(dim3 grid, dim3 block, F body) {
    #pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int by = 0; by < (int)grid.y; ++by)
        for (int bx = 0; bx < (int)grid.x; ++bx) {
            gridDim = grid; blockDim = block;
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0;
            for (unsigned ty = 0; ty < block.y; ++ty)
                for (unsigned tx = 0; tx < block.x; ++tx) {
                    threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = 0;
                    body();
                }
        }
}
#define REF_LAUNCH(grid, block, kernel, ...) rlpt_walk_grid(grid, block, [&] { kernel(__VA_ARGS__); })
#define REF_SYNC() 0
template <class T> static int ref_alloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, sizeof(T)); return *p ? 0 : 1; }
template <class T> static void ref_free(T* p) { free(p); }
static void ref_prefetch(void*, size_t) {}
/* utils/cuda_helpers.cu:16-25 is CUDA-runtime code; the host build seeds the shim generator the same way. */
static void init_rand_state(curandState* st, int width, int height) {
    int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= width || y >= height) return;
    curand_init(1984, x * height + y, 0, &st[x * height + y]);
}
struct RefTimer {
    std::chrono::steady_clock::time_point t0;
    void start() { t0 = std::chrono::steady_clock::now(); }
    double stop_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
#else
#include <cuda_runtime.h>
#include "cuda_helpers.cuh"
#define REF_LAUNCH(grid, block, kernel, ...) kernel<<<grid, block>>>(__VA_ARGS__)
#define REF_SYNC() ref_sync_check(__LINE__)
static int ref_sync_check(int line) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "ref_harness: CUDA error %d (%s) at line %d\n", (int)e, cudaGetErrorString(e), line); return 1; }
    return 0;
}
template <class T> static int ref_alloc(T** p, size_t n) {
    cudaError_t e = cudaMallocManaged((void**)p, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) { fprintf(stderr, "ref_harness: alloc failed: %s\n", cudaGetErrorString(e)); return 1; }
    memset((void*)*p, 0, (n ? n : 1) * sizeof(T));
    return 0;
}
template <class T> static void ref_free(T* p) { if (p) cudaFree((void*)p); }
static void ref_prefetch(void* p, size_t bytes) { if (p && bytes) cudaMemPrefetchAsync(p, bytes, 0, 0); }
struct RefTimer {
    cudaEvent_t a, b;
    RefTimer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~RefTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, 0); }
    double stop_ms() { cudaEventRecord(b, 0); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};
#endif

/* ---- harness-owned kernels (thin callers of reference device functions) ---- */
__global__ void k_ref_closest_hit(Scene* scene, const float* org, const float* dir, int n,
                                  int* type, int* index, float* dist, float* pos) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray ray(vec4(org[3 * i], org[3 * i + 1], org[3 * i + 2], 1.f), vec4(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], 1.f));
    ray.intersection.index = -1;
    ray.closest_intersection(scene);
    int t = (int)ray.intersection.intersection_type;   /* NOTHING 0, AREA_LIGHT 1, SURFACE 2 (rays/ray.cuh:30-34) */
    type[i] = t;
    index[i] = t ? ray.intersection.index : -1;
    dist[i] = ray.intersection.distance;
    pos[3 * i] = t ? ray.intersection.position.x : 0.f;
    pos[3 * i + 1] = t ? ray.intersection.position.y : 0.f;
    pos[3 * i + 2] = t ? ray.intersection.position.z : 0.f;
}
__global__ void k_ref_find_closest(RadianceMap* rm, const float* pos, const float* nrm, int n, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RadianceVolume* v = rm->find_closest_radiance_volume_iterative(MAX_DIST, vec4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 1.f),
                                                                   vec4(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], 1.f));
    out[i] = (int)(v - rm->radiance_volumes);
}
__global__ void k_ref_grid_dir(RadianceMap* rm, int vol, const float* gx, const float* gy, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RadianceVolume* v = &rm->radiance_volumes[vol];
    vec3 d = convert_grid_pos_to_direction(gx[i], gy[i], vec3(v->position), v->transformation_matrix);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
}

/* ---- state ---- */
static Scene* g_scene = nullptr;            /* managed (CUDA) / heap (host) */
static Camera* g_camera = nullptr;
static RadianceMap* g_rmap = nullptr;
static RadianceVolume* g_vols = nullptr;
static RadianceTreeElement* g_tree = nullptr;
static curandState* g_rand = nullptr;
static vec3* g_buffer = nullptr;
static int* g_path_lengths = nullptr;
static int* g_zero_contrib = nullptr;
static bool g_rand_ready = false;

static dim3 px_block() { return dim3(8, 8); }
static dim3 px_grid() { return dim3((SCREEN_WIDTH + 7) / 8, (SCREEN_HEIGHT + 7) / 8); }

static int ensure_frame_state() {
    if (!g_buffer) {
        if (ref_alloc(&g_buffer, (size_t)SCREEN_WIDTH * SCREEN_HEIGHT)) return 1;
        if (ref_alloc(&g_path_lengths, (size_t)SCREEN_WIDTH * SCREEN_HEIGHT)) return 1;
        if (ref_alloc(&g_zero_contrib, 1)) return 1;
        if (ref_alloc(&g_rand, (size_t)SCREEN_WIDTH * SCREEN_HEIGHT)) return 1;
    }
    if (!g_camera) {
        if (ref_alloc(&g_camera, 1)) return 1;
        Camera c(vec4(0.f, 0.f, -3.f, 1.f)); c.yaw_x = 0.f;   /* camera.cu:3-7 leaves yaw_x unset */
        memcpy((void*)g_camera, &c, sizeof(Camera));
    }
    if (!g_rand_ready) {
        REF_LAUNCH(px_grid(), px_block(), init_rand_state, g_rand, SCREEN_WIDTH, SCREEN_HEIGHT);
        if (REF_SYNC()) return 1;
        g_rand_ready = true;
    }
    return 0;
}

static int adopt_scene(Scene& h) {
    if (g_scene) { ref_free(g_scene->surfaces); ref_free(g_scene->area_lights); ref_free(g_scene->vertices); ref_free(g_scene); g_scene = nullptr; }
    if (ref_alloc(&g_scene, 1)) return 1;
    Surface* s; AreaLight* l; float* v;
    if (ref_alloc(&s, h.surfaces_count) || ref_alloc(&l, h.area_light_count) || ref_alloc(&v, h.vertices_count)) return 1;
    memcpy((void*)s, (void*)h.surfaces, sizeof(Surface) * h.surfaces_count);
    memcpy((void*)l, (void*)h.area_lights, sizeof(AreaLight) * h.area_light_count);
    memcpy((void*)v, (void*)h.vertices, sizeof(float) * h.vertices_count);
    memcpy((void*)g_scene, (void*)&h, sizeof(Scene));
    g_scene->surfaces = s; g_scene->area_lights = l; g_scene->vertices = v;
    return 0;
}

extern "C" {

int ref_dims(int* w, int* h, int* spp, int* bounces) {
    *w = SCREEN_WIDTH; *h = SCREEN_HEIGHT; *spp = SAMPLES_PER_PIXEL; *bounces = MAX_RAY_BOUNCES; return 0;
}
int ref_is_host(void) {
#ifdef RLPT_REF_HOST
    return 1;
#else
    return 0;
#endif
}
int ref_threads(void) {
#ifdef RLPT_REF_HOST
    return omp_get_max_threads();
#else
    return 0;
#endif
}

int ref_scene_cornell(void) {
    Scene h; h.load_cornell_box_scene();
    return adopt_scene(h);
}
int ref_scene_obj(const char* path, int lights_in_obj) {
    FILE* f = fopen(path, "r"); if (!f) return 2; fclose(f);
    Scene h; h.load_custom_scene(path, lights_in_obj != 0);
    return adopt_scene(h);
}
/* Scene from raw triangle arrays through the reference's own constructors (Surface/AreaLight/Material). */
int ref_scene_arrays(const float* sv, const float* srgb, int ns, const float* lv, const float* lrgb, int nl) {
    std::vector<Surface> S; std::vector<AreaLight> L; std::vector<float> V;
    for (int i = 0; i < ns; ++i) {
        const float* p = sv + 9 * i;
        Surface s(vec4(p[0], p[1], p[2], 1.f), vec4(p[3], p[4], p[5], 1.f), vec4(p[6], p[7], p[8], 1.f), Material(vec3(srgb[3 * i], srgb[3 * i + 1], srgb[3 * i + 2])));
        s.compute_and_set_normal(); S.push_back(s);
        for (int k = 0; k < 9; ++k) V.push_back(p[k]);
    }
    for (int i = 0; i < nl; ++i) {
        const float* p = lv + 9 * i;
        AreaLight a(vec4(p[0], p[1], p[2], 1.f), vec4(p[3], p[4], p[5], 1.f), vec4(p[6], p[7], p[8], 1.f), vec3(lrgb[3 * i], lrgb[3 * i + 1], lrgb[3 * i + 2]));
        a.compute_and_set_normal(); L.push_back(a);
        for (int k = 0; k < 9; ++k) V.push_back(p[k]);
    }
    Scene h; h.surfaces_count = ns; h.area_light_count = nl; h.vertices_count = (int)V.size();
    h.surfaces = S.data(); h.area_lights = L.data(); h.vertices = V.data();
    return adopt_scene(h);
}
/* RadianceVolume::read_radiance_volumes_to_surfaces (G/radiance_volumes/radiance_volume.cu:499-510): saved radiance volumes as
 * hemisphere geometry, the RENDER_SAVED_RADIANCE_VOLUMES path of Scene::load_custom_scene. Returns the surface count. */
int ref_saved_volumes_to_surfaces(const char* path, int max_surfaces, float* sv, float* srgb, float* snrm) {
#ifdef RLPT_REF_HOST
    std::vector<Surface> S;
    RadianceVolume::read_radiance_volumes_to_surfaces(std::string(path), S);
    for (int i = 0; i < (int)S.size() && i < max_surfaces; ++i) {
        Surface& s = S[i];
        float v[9] = { s.v0.x, s.v0.y, s.v0.z, s.v1.x, s.v1.y, s.v1.z, s.v2.x, s.v2.y, s.v2.z };
        memcpy(sv + 9 * i, v, sizeof v);
        srgb[3 * i] = s.material.diffuse_c.x; srgb[3 * i + 1] = s.material.diffuse_c.y; srgb[3 * i + 2] = s.material.diffuse_c.z;
        snrm[3 * i] = s.normal.x; snrm[3 * i + 1] = s.normal.y; snrm[3 * i + 2] = s.normal.z;
    }
    return (int)S.size();
#else
    return -1;
#endif
}
int ref_scene_counts(int* ns, int* nl) { if (!g_scene) return 1; *ns = g_scene->surfaces_count; *nl = g_scene->area_light_count; return 0; }
int ref_scene_get(float* sv, float* srgb, float* snrm, float* slum, float* lv, float* lrgb, float* lnrm, float* llum) {
    if (!g_scene) return 1;
    for (int i = 0; i < g_scene->surfaces_count; ++i) {
        Surface& s = g_scene->surfaces[i];
        float v[9] = { s.v0.x, s.v0.y, s.v0.z, s.v1.x, s.v1.y, s.v1.z, s.v2.x, s.v2.y, s.v2.z };
        memcpy(sv + 9 * i, v, sizeof v);
        srgb[3 * i] = s.material.diffuse_c.x; srgb[3 * i + 1] = s.material.diffuse_c.y; srgb[3 * i + 2] = s.material.diffuse_c.z;
        snrm[3 * i] = s.normal.x; snrm[3 * i + 1] = s.normal.y; snrm[3 * i + 2] = s.normal.z;
        slum[i] = s.material.luminance;
    }
    for (int i = 0; i < g_scene->area_light_count; ++i) {
        AreaLight& s = g_scene->area_lights[i];
        float v[9] = { s.v0.x, s.v0.y, s.v0.z, s.v1.x, s.v1.y, s.v1.z, s.v2.x, s.v2.y, s.v2.z };
        memcpy(lv + 9 * i, v, sizeof v);
        lrgb[3 * i] = s.diffuse_p.x; lrgb[3 * i + 1] = s.diffuse_p.y; lrgb[3 * i + 2] = s.diffuse_p.z;
        lnrm[3 * i] = s.normal.x; lnrm[3 * i + 1] = s.normal.y; lnrm[3 * i + 2] = s.normal.z;
        llum[i] = s.luminance;
    }
    return 0;
}
int ref_camera(float x, float y, float z, float yaw_y, float yaw_x) {
    if (!g_camera && ref_alloc(&g_camera, 1)) return 1;
    Camera c(vec4(x, y, z, 1.f)); c.yaw_y = yaw_y; c.yaw_x = yaw_x;
    memcpy((void*)g_camera, &c, sizeof(Camera));
    return 0;
}

int ref_closest_hit(const float* org, const float* dir, int n, int* type, int* index, float* dist, float* pos) {
    if (!g_scene) return 1;
    float *o, *d, *t, *p; int *ty, *ix;
    if (ref_alloc(&o, 3 * (size_t)n) || ref_alloc(&d, 3 * (size_t)n) || ref_alloc(&t, n) || ref_alloc(&p, 3 * (size_t)n) || ref_alloc(&ty, n) || ref_alloc(&ix, n)) return 1;
    memcpy(o, org, sizeof(float) * 3 * n); memcpy(d, dir, sizeof(float) * 3 * n);
    REF_LAUNCH(dim3((n + 63) / 64), dim3(64), k_ref_closest_hit, g_scene, o, d, n, ty, ix, t, p);
    int rc = REF_SYNC();
    memcpy(type, ty, sizeof(int) * n); memcpy(index, ix, sizeof(int) * n); memcpy(dist, t, sizeof(float) * n); memcpy(pos, p, sizeof(float) * 3 * n);
    ref_free(o); ref_free(d); ref_free(t); ref_free(p); ref_free(ty); ref_free(ix);
    return rc;
}

int ref_rmap_build(void) {
    if (!g_scene) return -1;
    srand(1);   /* the reference never seeds rand(); a fresh process starts from srand(1) (glibc) */
    std::vector<RadianceVolume> rvs; std::vector<RadianceTreeElement> tree;
    RadianceMap* hm = new RadianceMap(g_scene->surfaces, g_scene->surfaces_count, rvs, tree);
    ref_free(g_vols); ref_free(g_tree); ref_free(g_rmap);
    if (ref_alloc(&g_rmap, 1) || ref_alloc(&g_vols, rvs.size()) || ref_alloc(&g_tree, tree.size())) return -1;
    memcpy((void*)g_vols, (void*)rvs.data(), sizeof(RadianceVolume) * rvs.size());
    memcpy((void*)g_tree, (void*)tree.data(), sizeof(RadianceTreeElement) * tree.size());
    memcpy((void*)g_rmap, (void*)hm, sizeof(RadianceMap));
    g_rmap->radiance_volumes = g_vols; g_rmap->radiance_array = g_tree;
    g_rmap->radiance_volumes_count = (int)rvs.size(); g_rmap->radiance_array_size = (int)tree.size();
    return (int)rvs.size();
}
int ref_rmap_counts(int* nvol, int* ntree) { if (!g_rmap) return 1; *nvol = g_rmap->radiance_volumes_count; *ntree = g_rmap->radiance_array_size; return 0; }
int ref_rmap_get_volumes(float* pos, float* nrm, int* surf) {
    if (!g_rmap) return 1;
    for (int i = 0; i < g_rmap->radiance_volumes_count; ++i) {
        RadianceVolume& v = g_vols[i];
        pos[3 * i] = v.position.x; pos[3 * i + 1] = v.position.y; pos[3 * i + 2] = v.position.z;
        nrm[3 * i] = v.normal.x; nrm[3 * i + 1] = v.normal.y; nrm[3 * i + 2] = v.normal.z;
        surf[i] = (int)v.surface_index;
    }
    return 0;
}
int ref_rmap_get_tree(int* dim, int* leaf, unsigned* left, unsigned* right, float* data, float* pos, float* nrm) {
    if (!g_rmap) return 1;
    for (int i = 0; i < g_rmap->radiance_array_size; ++i) {
        RadianceTreeElement& e = g_tree[i];
        dim[i] = (int)e.dimension; leaf[i] = e.leaf ? 1 : 0; left[i] = e.left_idx; right[i] = e.right_idx; data[i] = e.data;
        pos[3 * i] = e.position.x; pos[3 * i + 1] = e.position.y; pos[3 * i + 2] = e.position.z;
        nrm[3 * i] = e.normal.x; nrm[3 * i + 1] = e.normal.y; nrm[3 * i + 2] = e.normal.z;
    }
    return 0;
}
int ref_rmap_get_state(float* q, float* cdf, unsigned* visits, float* irr) {
    if (!g_rmap) return 1;
    const int A = GRID_RESOLUTION * GRID_RESOLUTION;
    for (int i = 0; i < g_rmap->radiance_volumes_count; ++i) {
        RadianceVolume& v = g_vols[i];
        if (q) memcpy(q + (size_t)A * i, v.radiance_grid, sizeof(float) * A);
        if (cdf) memcpy(cdf + (size_t)A * i, v.radiance_distribution, sizeof(float) * A);
        if (visits) memcpy(visits + (size_t)A * i, v.visits, sizeof(unsigned) * A);
        if (irr) irr[i] = v.irradiance_accum;
    }
    return 0;
}
int ref_rmap_set_q(const float* q) {
    if (!g_rmap) return 1;
    const int A = GRID_RESOLUTION * GRID_RESOLUTION;
    for (int i = 0; i < g_rmap->radiance_volumes_count; ++i) memcpy(g_vols[i].radiance_grid, q + (size_t)A * i, sizeof(float) * A);
    return 0;
}
int ref_rmap_update_distributions(void) {
    if (!g_rmap) return 1;
    int n = g_rmap->radiance_volumes_count;
    REF_LAUNCH(dim3((n + 31) / 32), dim3(32), update_radiance_volume_distributions, g_rmap);
    return REF_SYNC();
}
int ref_find_closest(const float* pos, const float* nrm, int n, int* out, int on_device) {
    if (!g_rmap) return 1;
    if (!on_device) {   /* the function is __host__ __device__ in the reference: call it on the host */
        for (int i = 0; i < n; ++i) {
            RadianceVolume* v = g_rmap->find_closest_radiance_volume_iterative(MAX_DIST, vec4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 1.f),
                                                                               vec4(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], 1.f));
            out[i] = (int)(v - g_rmap->radiance_volumes);
        }
        return 0;
    }
    float *p, *m; int* o;
    if (ref_alloc(&p, 3 * (size_t)n) || ref_alloc(&m, 3 * (size_t)n) || ref_alloc(&o, n)) return 1;
    memcpy(p, pos, sizeof(float) * 3 * n); memcpy(m, nrm, sizeof(float) * 3 * n);
    REF_LAUNCH(dim3((n + 63) / 64), dim3(64), k_ref_find_closest, g_rmap, p, m, n, o);
    int rc = REF_SYNC();
    memcpy(out, o, sizeof(int) * n);
    ref_free(p); ref_free(m); ref_free(o);
    return rc;
}
int ref_grid_dir(int vol, const float* gx, const float* gy, int n, float* out) {
    if (!g_rmap) return 1;
    float *x, *y, *o;
    if (ref_alloc(&x, n) || ref_alloc(&y, n) || ref_alloc(&o, 3 * (size_t)n)) return 1;
    memcpy(x, gx, sizeof(float) * n); memcpy(y, gy, sizeof(float) * n);
    REF_LAUNCH(dim3((n + 63) / 64), dim3(64), k_ref_grid_dir, g_rmap, vol, x, y, n, o);
    int rc = REF_SYNC();
    memcpy(out, o, sizeof(float) * 3 * n);
    ref_free(x); ref_free(y); ref_free(o);
    return rc;
}

/* stats_per_frame: [avg_path_length, ms] per frame. out_mean: 3*W*H floats, x-major (pixel = x*H + y). */
int ref_render_default(int frames, float* out_mean, double* stats_per_frame) {
    if (!g_scene || ensure_frame_state()) return 1;
    const size_t np = (size_t)SCREEN_WIDTH * SCREEN_HEIGHT;
    std::vector<double> acc(3 * np, 0.0);
    RefTimer tm;
    ref_prefetch(g_scene->surfaces, sizeof(Surface) * g_scene->surfaces_count);
    for (int f = 0; f < frames; ++f) {
        tm.start();
        REF_LAUNCH(px_grid(), px_block(), draw_default_path_tracing, g_buffer, g_rand, g_camera, g_scene, g_path_lengths);
        double ms = tm.stop_ms();
        if (REF_SYNC()) return 1;
        long long total = 0;
        for (size_t i = 0; i < np; ++i) total += g_path_lengths[i];
        for (size_t i = 0; i < np; ++i) { acc[3 * i] += g_buffer[i].x; acc[3 * i + 1] += g_buffer[i].y; acc[3 * i + 2] += g_buffer[i].z; }
        if (stats_per_frame) { stats_per_frame[2 * f] = (double)total / (double)np; stats_per_frame[2 * f + 1] = ms; }
        memset((void*)g_buffer, 0, sizeof(vec3) * np);
    }
    for (size_t i = 0; i < 3 * np; ++i) out_mean[i] = (float)(acc[i] / frames);
    return 0;
}

/* stats_per_frame: [avg_path_length, zero_contribution_paths, ms_trace, ms_update, nan_pixels] per frame.
 * out_mean: NaN-guarded mean over frames >= skip_frames; out_last: the last frame as the reference would save it. */
int ref_render_sarsa(int frames, int skip_frames, float* out_mean, float* out_last, double* stats_per_frame) {
    if (!g_scene || !g_rmap || ensure_frame_state()) return 1;
    const size_t np = (size_t)SCREEN_WIDTH * SCREEN_HEIGHT;
    std::vector<double> acc(3 * np, 0.0); std::vector<int> cnt(np, 0);
    RefTimer tm;
    int nv = g_rmap->radiance_volumes_count;
    ref_prefetch(g_vols, sizeof(RadianceVolume) * nv);
    ref_prefetch(g_tree, sizeof(RadianceTreeElement) * g_rmap->radiance_array_size);
    for (int f = 0; f < frames; ++f) {
        *g_zero_contrib = 0;
        tm.start();
        REF_LAUNCH(px_grid(), px_block(), draw_reinforcement_path_tracing, g_buffer, g_rand, g_rmap, g_camera, g_scene, g_path_lengths, g_zero_contrib);
        double ms_trace = tm.stop_ms();
        if (REF_SYNC()) return 1;
        tm.start();
        REF_LAUNCH(dim3((nv + 31) / 32), dim3(32), update_radiance_volume_distributions, g_rmap);
        double ms_upd = tm.stop_ms();
        if (REF_SYNC()) return 1;
        long long total = 0; int nan_px = 0;
        for (size_t i = 0; i < np; ++i) total += g_path_lengths[i];
        for (size_t i = 0; i < np; ++i) {
            vec3 c = g_buffer[i];
            bool bad = !(std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z));
            if (bad) { nan_px++; continue; }
            if (f >= skip_frames) { acc[3 * i] += c.x; acc[3 * i + 1] += c.y; acc[3 * i + 2] += c.z; cnt[i]++; }
        }
        if (out_last && f == frames - 1) for (size_t i = 0; i < np; ++i) { out_last[3 * i] = g_buffer[i].x; out_last[3 * i + 1] = g_buffer[i].y; out_last[3 * i + 2] = g_buffer[i].z; }
        if (stats_per_frame) {
            double* s = stats_per_frame + 5 * f;
            s[0] = (double)total / (double)np; s[1] = (double)*g_zero_contrib; s[2] = ms_trace; s[3] = ms_upd; s[4] = nan_px;
        }
        memset((void*)g_buffer, 0, sizeof(vec3) * np);
    }
    if (out_mean) for (size_t i = 0; i < np; ++i) for (int c = 0; c < 3; ++c) out_mean[3 * i + c] = cnt[i] ? (float)(acc[3 * i + c] / cnt[i]) : 0.f;
    return 0;
}

}  /* extern "C" */
