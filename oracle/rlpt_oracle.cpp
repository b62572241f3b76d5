/* oracle/rlpt_oracle.cpp -- TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's reinforcement-learned
 * path tracing hot path (GPU_Rendering_Engine/Source, shorthand G/). It is the checker the tests and
 * bench.py's cpu_baseline leg compare the CUDA product against; the product never includes, links or calls it.
 *
 * Parity status: PINNED. Every function below is checked (tests/test_oracle_vs_reference.py) against the
 * reference itself -- oracle/_ref/libref_host.so, the unmodified reference sources compiled for the host --
 * and against the committed vectors in tests/golden/ generated from that build (tests/golden/make_golden.py).
 * On the GPU box the same functions are checked against oracle/_ref/libref_cuda.so (the reference's kernels).
 *
 * Arithmetic: compiled with -ffp-contract=off, so every operation here is an individually rounded IEEE op and
 * fmaf() appears only where it is written. Two modes for the two places where parity is bit-exact:
 *   fma_mode 0  glm's expression order with no contraction  == the reference built for the host (libref_host.so)
 *   fma_mode 1  the contraction pattern nvcc 12.9 emits for sm_100a (read from the SASS of G/rays/ray.cu and
 *               G/radiance_volumes/radiance_map.cu; DESIGN.md section "Bit-exact arithmetic") == libref_cuda.so
 * Random numbers: Philox4x32-10, key (seed, 0), counter (pixel, sample, bounce, purpose) -- the product's
 * counter-based scheme, so product and oracle trace the same paths (the reference's XORWOW stream is
 * not reproduced; the reference is compared statistically).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int GRID = 12;                 /* G/constants/radiance_volumes_settings.h:9 */
constexpr int A = GRID * GRID;
constexpr float PI_F = (float)M_PI;
constexpr float RHO = 1.f / (2.f * 3.1415926535f);           /* G/constants/image_settings.h:13 */
constexpr float GRID_RHO = 1.f / ((float)GRID * (float)GRID); /* radiance_volumes_settings.h:10 */

struct V3 { float x, y, z; };
static inline V3 sub(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static inline float dot_plain(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }   /* glm compute_dot<vec3>: tmp.x + tmp.y + tmp.z */
static inline V3 cross(V3 x, V3 y) { return { x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y }; } /* glm func_geometric.inl */
/* glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt (glm/detail/func_geometric.inl:82-90).
 * fma_mode 1: dot contracts to fma(z,z,fma(x,x,y*y)) (SASS of Ray::Ray, G/rays/ray.cu:6-14). */
static inline V3 normalize(V3 v, int fma_mode) {
    float d = fma_mode ? fmaf(v.z, v.z, fmaf(v.x, v.x, v.y * v.y)) : dot_plain(v, v);
    float inv = 1.f / sqrtf(d);
    return { v.x * inv, v.y * inv, v.z * inv };
}

struct Tri { V3 v0, v1, v2, n; };
struct Scene {
    std::vector<Tri> surf, light;
    std::vector<V3> surf_rgb, light_rgb;
    std::vector<float> surf_lum, light_lum;
} g_scene;

static float luminance(V3 c) {            /* G/objects/material.cu:4-14, G/lights/area_light.cu:13-21 */
    float mx = std::max(c.z, std::max(c.x, c.y)), mn = std::min(c.z, std::min(c.x, c.y));
    return 0.5f * (mx + mn);
}
static V3 tri_normal(const Tri& t) {      /* G/objects/triangle.cu:67-76: normalize(cross(e2, e1)), host code */
    V3 e1 = sub(t.v1, t.v0), e2 = sub(t.v2, t.v0);
    return normalize(cross(e2, e1), 0);
}

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al. 2011) */
static inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
/* cuRAND's curand_uniform convention, (0,1]: x*2^-32 + 2^-33 (the reference draws with curand_uniform everywhere). */
static inline float u01(uint32_t x) { return x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }
enum { PURPOSE_CAMERA = 0, PURPOSE_BOUNCE = 1 };
static inline void draw4(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t purpose, float u[4]) {
    uint32_t c[4] = { pixel, sample, bounce, purpose };
    philox4x32_10(seed, 0u, c);
    for (int i = 0; i < 4; ++i) u[i] = u01(c[i]);
}

/* ------------------------------------------------------------------ closest hit: G/rays/ray.cu:16-141 */
struct Hit { int type; int index; float t; V3 pos; V3 n; };   /* type: 0 NOTHING, 1 AREA_LIGHT, 2 SURFACE (G/rays/ray.cuh:30-34) */

/* One Ray::intersects + Ray::cramer (ray.cu:39-74,115-141). d is the ray direction already scaled by
 * SCREEN_HEIGHT (ray.cu:53). Returns true and t when the triangle is hit (t,u,v >= 0, u+v <= 1, no epsilon). */
static inline bool tri_test(const Tri& tr, V3 o, V3 d, int fma_mode, float& t_out) {
    V3 e1 = sub(tr.v1, tr.v0), e2 = sub(tr.v2, tr.v0), b = sub(o, tr.v0);
    float a0 = 0.f - d.x, a1 = 0.f - d.y, a2 = 0.f - d.z;      /* A[0] = -dir */
    float detA, dx, dy, dz;
    if (!fma_mode) {
        /* glm determinant(mat3) (glm/detail/func_matrix.inl:211-220), m[col][row], columns (-dir, e1, e2) */
        auto det3 = [](V3 c0, V3 c1, V3 c2) {
            return (c0.x * (c1.y * c2.z - c2.y * c1.z) - c1.x * (c0.y * c2.z - c2.y * c0.z)) + c2.x * (c0.y * c1.z - c1.y * c0.z);
        };
        V3 m0 = { a0, a1, a2 };
        detA = det3(m0, e1, e2);
        if (!(detA != 0.f)) return false;
        dx = det3(b, e1, e2); dy = det3(m0, b, e2); dz = det3(m0, e1, b);
    } else {
        /* nvcc 12.9 -fmad=true, sm_100a: every 2x2 minor is fma(p,q,-(r*s)) with r*s rounded first, except the two
         * minors shared between the x and z systems, which are a difference of two rounded products; each 3x3 is
         * fma(c2x, M3, fma(c0x, M1, -(c1x*M2))).                                  (SASS, Ray::intersects) */
        float T1 = fmaf(e1.y, e2.z, -(e1.z * e2.y));
        float T2 = fmaf(e2.z, a1, -(e2.y * a2));
        float T3 = fmaf(e1.z, a1, -(e1.y * a2));
        detA = fmaf(e2.x, T3, fmaf(a0, T1, -(e1.x * T2)));
        if (!(detA != 0.f)) return false;
        float p75 = e1.z * b.y, p80 = e1.y * b.z;
        float U2 = fmaf(e2.z, b.y, -(e2.y * b.z));
        float U3 = p75 - p80;
        dx = fmaf(e2.x, U3, fmaf(T1, b.x, -(e1.x * U2)));
        float V3_ = fmaf(a1, b.z, -(a2 * b.y));
        dy = fmaf(e2.x, V3_, fmaf(a0, U2, -(T2 * b.x)));
        float W1 = p80 - p75;
        dz = fmaf(T3, b.x, fmaf(a0, W1, -(e1.x * V3_)));
    }
    float t = dx / detA, u = dy / detA, v = dz / detA;
    if (t >= 0.f && u >= 0.f && v >= 0.f && u + v <= 1.f) { t_out = t; return true; }
    return false;
}

/* Ray::closest_intersection (ray.cu:16-36): surfaces first, then lights, strict < on distance. `dir` must already be
 * normalised the way Ray::Ray does it. H = SCREEN_HEIGHT. */
static Hit closest_hit(V3 o, V3 dir, float H, int fma_mode) {
    Hit h; h.type = 0; h.index = -1; h.t = 999999.f; h.pos = { 0, 0, 0 }; h.n = { 0, 0, 0 };
    V3 d = { dir.x * H, dir.y * H, dir.z * H };
    for (int kind = 0; kind < 2; ++kind) {
        const std::vector<Tri>& tris = kind == 0 ? g_scene.surf : g_scene.light;
        for (int i = 0; i < (int)tris.size(); ++i) {
            float t;
            if (tri_test(tris[i], o, d, fma_mode, t) && t < h.t) {
                h.t = t; h.index = i; h.type = kind == 0 ? 2 : 1; h.n = tris[i].n;
                /* position = start + t*dir (ray.cu:65); nvcc contracts to fma */
                if (fma_mode) h.pos = { fmaf(d.x, t, o.x), fmaf(d.y, t, o.y), fmaf(d.z, t, o.z) };
                else h.pos = { o.x + t * d.x, o.y + t * d.y, o.z + t * d.z };
            }
        }
    }
    return h;
}

/* ------------------------------------------------------------------ hemisphere helpers: G/utils/hemisphere_helpers.cu */
/* map(): Shirley-Chiu square -> hemisphere (hemisphere_helpers.cu:134-226). Offsets are double expressions
 * assigned to float; phi is evaluated in double and rounded to float. */
static void sc_map(float x, float y, float& xr, float& yr, float& zr) {
    float xx, yy, offset;
    x = 2 * x - 1; y = 2 * y - 1;
    if (y > -x) {
        if (y < x) { xx = x; if (y > 0) { offset = 0; yy = y; } else { offset = (float)((7 * M_PI) / 4); yy = x + y; } }
        else { xx = y; if (x > 0) { offset = (float)(M_PI / 4); yy = y - x; } else { offset = (float)((2 * M_PI) / 4); yy = -x; } }
    } else {
        if (y > x) { xx = -x; if (y > 0) { offset = (float)((3 * M_PI) / 4); yy = -x - y; } else { offset = (float)((4 * M_PI) / 4); yy = -y; } }
        else {
            xx = -y;
            if (x > 0) { offset = (float)((6 * M_PI) / 4); yy = x; }
            else if (y != 0) { offset = (float)((5 * M_PI) / 4); yy = x - y; }
            else { xr = 0.f; yr = 1.f; zr = 0.f; return; }
        }
    }
    float theta = acosf(1 - xx * xx);
    float phi = (float)((double)offset + (M_PI / 4) * (double)(yy / xx));
    xr = sinf(theta) * cosf(phi); yr = cosf(theta); zr = sinf(theta) * sinf(phi);
}
/* create_normal_coordinate_system (hemisphere_helpers.cu:31-44) */
static void tangent_frame(V3 n, V3& T, V3& B) {
    if (fabsf(n.x) > fabsf(n.y)) T = normalize(V3{ n.z, 0.f, -n.x }, 0);
    else T = normalize(V3{ 0.f, -n.z, n.y }, 0);
    B = cross(n, T);
}
/* convert_grid_pos_to_direction (hemisphere_helpers.cu:96-105) with the matrix of create_transformation_matrix
 * (:48-63) = columns (T, N, B, position): world = T*xh + N*yh + B*zh + pos; dir = normalize(world - pos). */
static V3 grid_dir(float gx, float gy, V3 pos, V3 n) {
    float xh, yh, zh; sc_map(gx / (float)GRID, gy / (float)GRID, xh, yh, zh);
    V3 T, B; tangent_frame(n, T, B);
    V3 w = { ((T.x * xh + n.x * yh) + B.x * zh) + pos.x, ((T.y * xh + n.y * yh) + B.y * zh) + pos.y, ((T.z * xh + n.z * yh) + B.z * zh) + pos.z };
    return normalize(sub(w, pos), 0);
}
/* sample_random_direction_around_intersection + uniform_hemisphere_sample (hemisphere_helpers.cu:8-25,67-93) */
static V3 uniform_hemisphere_dir(V3 n, float r1, float r2) {
    V3 T, B; tangent_frame(n, T, B);
    float sin_theta = sqrtf(1 - r1 * r1);
    float phi = (float)(2 * M_PI * (double)r2);
    float sx = sin_theta * cosf(phi), sy = r1, sz = sin_theta * sinf(phi);
    return { (sx * B.x + sy * n.x) + sz * T.x, (sx * B.y + sy * n.y) + sz * T.y, (sx * B.z + sy * n.z) + sz * T.z };
}

/* ------------------------------------------------------------------ radiance volumes + kd-tree */
struct Volume { V3 pos; V3 n; int surface; };
struct TreeEl { int dim; int leaf; unsigned left, right; float data; V3 pos; V3 n; };   /* G/radiance_volumes/radiance_tree.cuh:19-27 */
struct RMap {
    std::vector<Volume> vol;
    std::vector<TreeEl> tree;
    std::vector<float> q, cdf, irr;       /* nvol*A, nvol*A, nvol */
    std::vector<unsigned> visits;
    std::vector<double> acc_sum; std::vector<unsigned> acc_cnt;   /* batched-TD accumulators (td_mode 1) */
} g_rm;

/* Triangle::compute_area (G/objects/triangle.cu:17-26), host arithmetic: pow(float,int) and sqrt run in double */
static float tri_area(const Tri& t) {
    V3 a = sub(t.v1, t.v0), b = sub(t.v2, t.v0);
    float e = sqrtf(dot_plain(a, a)) * sqrtf(dot_plain(b, b));
    float c = dot_plain(a, b) / e;
    float s = (float)sqrt(1 - pow((double)c, 2));
    return 0.5f * e * s;
}
/* Triangle::sample_position_on_plane (triangle.cu:30-45): host rand(), rejection a1+a2 <= 1 */
static V3 sample_on_triangle(const Tri& t) {
    float a1, a2; V3 p;
    do {
        a1 = (float)rand() / (float)RAND_MAX; a2 = (float)rand() / (float)RAND_MAX;
        V3 e1 = sub(t.v1, t.v0), e2 = sub(t.v2, t.v0);
        p = { (t.v0.x + a1 * e1.x) + a2 * e2.x, (t.v0.y + a1 * e1.y) + a2 * e2.y, (t.v0.z + a1 * e1.z) + a2 * e2.z };
    } while (a1 + a2 > 1.f);
    return p;
}

struct KdBuild {   /* RadianceTree (G/radiance_volumes/radiance_tree.cu:12-62): sort on dim, median split, X->Y->Z */
    int dim = 0; float median = 0.f; int volume = -1; KdBuild* l = nullptr; KdBuild* r = nullptr;
    ~KdBuild() { delete l; delete r; }
};
static KdBuild* kd_build(std::vector<int>& ids, int dim) {
    KdBuild* node = new KdBuild; node->dim = dim;
    auto coord = [&](int id) { const V3& p = g_rm.vol[id].pos; return dim == 0 ? p.x : (dim == 1 ? p.y : p.z); };
    int n = (int)ids.size();
    if (n == 0) return node;
    if (n == 1) { node->median = coord(ids[0]); node->volume = ids[0]; return node; }
    std::sort(ids.begin(), ids.end(), [&](int a, int b) { return coord(a) < coord(b); });   /* std::sort, same comparator outcome as sort_on_x/y/z */
    int mi;
    if (n % 2 == 0) { mi = n / 2 - 1; node->median = (coord(ids[mi]) + coord(ids[mi + 1])) / 2; }
    else { mi = n / 2; node->median = coord(ids[mi]); }
    std::vector<int> L(ids.begin(), ids.begin() + mi + 1), R(ids.begin() + mi + 1, ids.end());
    int nd = (dim + 1) % 3;
    node->l = kd_build(L, nd); node->r = kd_build(R, nd);
    return node;
}
/* convert_to_array / traverse_and_insert (radiance_tree.cu:135-196): children appended in pairs when the parent is visited */
static void kd_flatten(KdBuild* t, std::vector<TreeEl>& out, int idx) {
    int last = (int)out.size() - 1;
    if (t->volume >= 0) {
        TreeEl e; e.dim = t->dim; e.leaf = 1; e.left = e.right = 0; e.data = (float)t->volume; e.pos = g_rm.vol[t->volume].pos; e.n = g_rm.vol[t->volume].n;
        out[idx] = e; return;
    }
    out[idx].left = last + 1; out[idx].right = last + 2;
    int nd = (t->dim + 1) % 3;
    TreeEl a; a.dim = nd; a.leaf = 0; a.left = a.right = 0; a.data = t->l->median; a.pos = { 0, 0, 0 }; a.n = { 0, 0, 0 };
    TreeEl b = a; b.data = t->r->median;
    out.push_back(a); out.push_back(b);
    kd_flatten(t->l, out, last + 1); kd_flatten(t->r, out, last + 2);
}

/* RadianceMap::find_closest_radiance_volume_iterative (G/radiance_volumes/radiance_map.cu:150-203) with Stack
 * (G/utils/stack.cu). Never returns "none": starts from volume 0 at distance |p - tree[0].position|. */
static int kd_find(V3 p, V3 n, float max_dist, int fma_mode) {
    const std::vector<TreeEl>& T = g_rm.tree;
    auto dist = [&](V3 a, V3 b) {
        V3 d = sub(a, b);
        float s = fma_mode ? fmaf(d.z, d.z, fmaf(d.x, d.x, d.y * d.y)) : dot_plain(d, d);
        return sqrtf(s);
    };
    int stack[64]; int top = 0; const int cap = (int)T.size();
    auto push = [&](int v) { if (top < cap - 1 && top < 64) stack[top++] = v; };
    push(0);
    int best = 0; float best_d = dist(p, T[0].pos);
    while (top > 0) {
        int idx = stack[--top];
        const TreeEl& e = T[idx];
        if (e.leaf) {
            float d = dist(e.pos, p);
            if (n.x == e.n.x && n.y == e.n.y && n.z == e.n.z && d < best_d) { best = (int)e.data; best_d = d; }
        } else {
            float c = e.dim == 0 ? p.x : (e.dim == 1 ? p.y : p.z);
            float delta = c - e.data;
            bool within = (double)delta * (double)delta < (double)max_dist;   /* pow(delta,2) < max_dist, evaluated in double */
            if (delta < 0) { if (within) push((int)e.right); push((int)e.left); }
            else { if (within) push((int)e.left); push((int)e.right); }
        }
    }
    return best;
}

/* per-cell cos(theta) at cell centres, as update_radiance_distribution evaluates it (radiance_volume.cu:156-158) */
static float cell_cos(const Volume& v, int x, int y) {
    V3 d = grid_dir((float)x + 0.5f, (float)y + 0.5f, v.pos, v.n);
    return dot_plain(d, v.n);
}
/* RadianceVolume::update_radiance_distribution (G/radiance_volumes/radiance_volume.cu:149-188) */
static void update_distribution(const Volume& v, const float* q, float* cdf) {
    float total = 0.0000000001f;
    float temp[A];
    for (int x = 0; x < GRID; ++x) for (int y = 0; y < GRID; ++y) {
        float t = q[x * GRID + y] * cell_cos(v, x, y);
        t = t > 0.f ? t : 0.f;     /* DISTRIBUTION_THRESHOLD 0 */
        temp[x * GRID + y] = t; total += t;
    }
    float prev = 0.f;
    for (int k = 0; k < A; ++k) { float r = temp[k] / total + prev; cdf[k] = r; prev = r; }
}
/* initial state: RadianceVolume ctor (radiance_volume.cu:12-24,49-89). irradiance_accum is evaluated there with
 * whatever frame the temporary happens to hold; the well-defined value it converges to -- the
 * cell-centre sum -- is what is restated. */
static void init_volume_state(int i) {
    const Volume& v = g_rm.vol[i];
    const float initial = (1.f / ((float)GRID * (float)GRID)) * 100.f;
    float lum = g_scene.surf_lum[v.surface];
    float irr = 0.f;
    for (int x = 0; x < GRID; ++x) for (int y = 0; y < GRID; ++y) {
        int k = x * GRID + y;
        g_rm.q[(size_t)i * A + k] = initial;
        g_rm.cdf[(size_t)i * A + k] = k * (1.f / ((float)GRID * (float)GRID));
        g_rm.visits[(size_t)i * A + k] = 0;
        irr += cell_cos(v, x, y) * (float)(lum / M_PI) * initial;
    }
    g_rm.irr[i] = irr;
}

/* RadianceVolume::sample_direction_from_radiance_distribution (radiance_volume.cu:192-244): returns the sector, or -1
 * when r falls past the last bin (the reference then returns direction 0, pdf 0). */
static int sample_sector(const float* cdf, float r, float& pdf) {
    if (r <= cdf[0]) { pdf = RHO * (cdf[0] / GRID_RHO); return 0; }
    int start = 0, end = A - 1;
    while (start <= end) {
        int mid = (end + start) / 2;
        float mv = cdf[mid], pv = mid > 0 ? cdf[mid - 1] : cdf[0];   /* mid == 0 is unreachable here unless r > cdf[0] fails above */
        if (r < mv && pv <= r) { pdf = RHO * ((mv - pv) / GRID_RHO); return mid; }
        else if (mv < r) start = mid + 1;
        else end = mid - 1;
    }
    pdf = 0.f; return -1;
}

/* Product behaviour where the reference's search fails (r past the last bin, or r exactly equal to a CDF entry the
 * binary search probes): the first bin k >= 1 with cdf[k] > r, else the last bin of non-zero width. Stated deviation. */
static int sample_sector_fallback(const float* cdf, float r, float& pdf) {
    int sector = 1; while (sector < A && !(cdf[sector] > r)) ++sector;
    if (sector >= A) { sector = A - 1; while (sector > 0 && !(cdf[sector] - cdf[sector - 1] > 0.f)) --sector; }
    float pv = sector > 0 ? cdf[sector - 1] : 0.f; pdf = RHO * ((cdf[sector] - pv) / GRID_RHO);
    return sector;
}

/* ------------------------------------------------------------------ path tracers */
struct Cfg {
    int width, height, spp, max_bounces; float env; uint32_t seed; float cam[3]; float yaw_y, yaw_x;
    int fma_mode; int td_mode; int clamp_last_bin; float max_dist;
};
/* Ray::sample_ray_through_pixel + rotate_ray (ray.cu:144-172) */
static void camera_ray(const Cfg& c, int px, int py, float u0, float u1, V3& o, V3& d) {
    float x = (float)px + u0, y = (float)py + u1;
    V3 dir = { x - (float)c.width / 2.f, y - (float)c.height / 2.f, (float)c.height };
    dir = normalize(dir, c.fma_mode);
    float cy = cosf(c.yaw_y), sy = sinf(c.yaw_y), cx = cosf(c.yaw_x), sx = sinf(c.yaw_x);
    /* R[0]=(cos,0,sin,0), R[1]=(0,1,0,0), R[2]=(-sin,0,cos,0), R[3]=(0,0,0,1); direction.w = 1 */
    V3 r1 = { (cy * dir.x + 0.f * dir.y) + (-sy) * dir.z, dir.y, (sy * dir.x + 0.f * dir.y) + cy * dir.z };
    V3 r2 = { r1.x, (cx * r1.y) + sx * r1.z, ((-sx) * r1.y) + cx * r1.z };
    if (c.yaw_y == 0.f && c.yaw_x == 0.f) r2 = dir;
    o = { c.cam[0], c.cam[1], c.cam[2] }; d = r2;
}

/* path_trace_iterative (G/path_tracing/default_path_tracing.cu:36-88) */
static V3 trace_default(const Cfg& c, uint32_t pixel, int px, int py, uint32_t sample, int& len) {
    float u[4]; draw4(c.seed, pixel, sample, 0, PURPOSE_CAMERA, u);
    V3 o, d; camera_ray(c, px, py, u[0], u[1], o, d);
    V3 thr = { 1.f, 1.f, 1.f };
    for (int i = 0; i < c.max_bounces; ++i) {
        Hit h = closest_hit(o, d, (float)c.height, c.fma_mode);
        if (h.type == 0) { len = i + 1; return { thr.x * c.env, thr.y * c.env, thr.z * c.env }; }
        if (h.type == 1) { len = i + 1; V3 p = g_scene.light_rgb[h.index]; return { thr.x * p.x, thr.y * p.y, thr.z * p.z }; }
        draw4(c.seed, pixel, sample, (uint32_t)i, PURPOSE_BOUNCE, u);
        float cos_theta = u[0];
        V3 nd = uniform_hemisphere_dir(h.n, u[0], u[1]);
        V3 rgb = g_scene.surf_rgb[h.index];
        V3 brdf = { rgb.x / PI_F, rgb.y / PI_F, rgb.z / PI_F };
        thr = { (thr.x * brdf.x * cos_theta) / RHO, (thr.y * brdf.y * cos_theta) / RHO, (thr.z * brdf.z * cos_theta) / RHO };
        o = { h.pos.x + 0.00001f * nd.x, h.pos.y + 0.00001f * nd.y, h.pos.z + 0.00001f * nd.z };
        d = normalize(nd, c.fma_mode);
    }
    len = c.max_bounces; return { 0.f, 0.f, 0.f };
}

/* RadianceVolume::temporal_difference_update + expected_sarsa_irradiance (radiance_volume.cu:283-301,94-112).
 * td_mode 0: live update as the reference does it (racy there, serialised here).
 * td_mode 1: accumulate (sum of targets, count); merged at frame end by merge_frame() -- the product's scheme. */
static void td_update(int vol, int sector, float target, int td_mode) {
    size_t k = (size_t)vol * A + sector;
    if (td_mode == 1) {
        #pragma omp atomic
        g_rm.acc_sum[k] += (double)target;
        #pragma omp atomic
        g_rm.acc_cnt[k] += 1u;
        return;
    }
    const float thresh = (1.f / ((float)GRID * (float)GRID)) * 0.8f;
    #pragma omp critical(rlpt_td)
    {
        unsigned vs = g_rm.visits[k];
        float alpha = 1.f / (1.f + (float)vs);
        float upd = ((1.f - alpha) * g_rm.q[k]) + (alpha * target);
        upd = upd > thresh ? upd : thresh;
        g_rm.visits[k] = vs + 1;
        const Volume& v = g_rm.vol[vol];
        int sx = sector / GRID, sy = sector % GRID;
        V3 dir = grid_dir((float)sx, (float)sy, v.pos, v.n);          /* corner of the cell, radiance_volume.cu:97 */
        float ct = dot_plain(dir, v.n);
        float brdf = g_scene.surf_lum[v.surface] / PI_F;
        g_rm.irr[vol] = (g_rm.irr[vol] - (g_rm.q[k] * ct * brdf)) + (upd * ct * brdf);
        g_rm.q[k] = upd;
    }
}
/* frame-end merge of the batched accumulators: running mean over all targets seen so far, clamp, then the
 * irradiance estimate recomputed from cell centres (the product's stated deviations). */
static void merge_frame() {
    const float thresh = (1.f / ((float)GRID * (float)GRID)) * 0.8f;
    int nv = (int)g_rm.vol.size();
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nv; ++i) {
        const Volume& v = g_rm.vol[i];
        float irr = 0.f; float brdf = g_scene.surf_lum[v.surface] / PI_F;
        for (int k = 0; k < A; ++k) {
            size_t j = (size_t)i * A + k;
            unsigned c = g_rm.acc_cnt[j];
            if (c) {
                float vs = (float)g_rm.visits[j];
                float qn = (vs * g_rm.q[j] + (float)g_rm.acc_sum[j]) / (vs + (float)c);
                g_rm.q[j] = qn > thresh ? qn : thresh;
                g_rm.visits[j] += c; g_rm.acc_cnt[j] = 0; g_rm.acc_sum[j] = 0.0;
            }
            irr += g_rm.q[j] * cell_cos(v, k / GRID, k % GRID) * brdf;
        }
        g_rm.irr[i] = irr;
    }
}

static int g_max_direction = 0;      /* orc_set_max_direction: greedy debug sampling (the thesis' 1-spp figures) */
/* path_trace_reinforcement_iterative (G/path_tracing/reinforcement_path_tracing.cu:48-120) +
 * RadianceMap::temporal_difference_update_radiance_volume_sector (radiance_map.cu:111-146) */
static V3 trace_sarsa(const Cfg& c, uint32_t pixel, int px, int py, uint32_t sample, int& len, bool& failed) {
    float u[4]; draw4(c.seed, pixel, sample, 0, PURPOSE_CAMERA, u);
    V3 o, d; camera_ray(c, px, py, u[0], u[1], o, d);
    V3 thr = { 1.f, 1.f, 1.f };
    int cur_vol = -1, cur_sector = -1; float cur_brdf = 0.f;
    failed = false;
    for (int i = 0; i < c.max_bounces; ++i) {
        Hit h = closest_hit(o, d, (float)c.height, c.fma_mode);
        if (i > 0) {
            if (cur_vol >= 0 && cur_sector != -1) {
                if (h.type == 0) { td_update(cur_vol, cur_sector, cur_brdf * c.env, c.td_mode); cur_vol = -1; }
                else if (h.type == 1) { td_update(cur_vol, cur_sector, cur_brdf * g_scene.light_lum[h.index], c.td_mode); cur_vol = -1; }
                else {
                    int nv = kd_find(h.pos, h.n, c.max_dist, c.fma_mode);
                    float est = g_rm.irr[nv] * ((2.f * PI_F) / ((float)(GRID * GRID)));    /* get_irradiance_estimate, radiance_volume.cu:305-307 */
                    td_update(cur_vol, cur_sector, est * cur_brdf, c.td_mode);
                    cur_vol = nv;
                }
                cur_sector = -1;
            }
        } else if (h.type == 2) cur_vol = kd_find(h.pos, h.n, c.max_dist, c.fma_mode);
        if (h.type == 0) { len = i + 1; return { thr.x * c.env, thr.y * c.env, thr.z * c.env }; }
        if (h.type == 1) { len = i + 1; V3 p = g_scene.light_rgb[h.index]; return { thr.x * p.x, thr.y * p.y, thr.z * p.z }; }
        draw4(c.seed, pixel, sample, (uint32_t)i, PURPOSE_BOUNCE, u);
        const float* cdf = &g_rm.cdf[(size_t)cur_vol * A];
        float pdf = 0.f;
        int sector;
        if (g_max_direction) {
            /* RadianceVolume::sample_max_direction_from_radiance_distribution (radiance_volume.cu:248-278): the first cell holding the largest Q;
             * pdf from the width of its CDF bin. (The reference takes cdf[0] - cdf[0] = 0 for cell 0; the proper width cdf[0] is used here, the same
             * repair as the clamped last bin.) */
            const float* q = &g_rm.q[(size_t)cur_vol * A];
            sector = 0; float mx = q[0];
            for (int k = 0; k < A; ++k) if (mx < q[k]) { mx = q[k]; sector = k; }
            pdf = RHO * ((cdf[sector] - (sector ? cdf[sector - 1] : 0.f)) / GRID_RHO);
        } else {
            sector = sample_sector(cdf, u[0], pdf);
            if (sector < 0 && c.clamp_last_bin) sector = sample_sector_fallback(cdf, u[0], pdf);
        }
        V3 nd;
        if (sector < 0) { failed = true; nd = { 0.f, 0.f, 0.f }; }
        else {
            cur_sector = sector;
            const Volume& v = g_rm.vol[cur_vol];
            nd = grid_dir((float)(sector / GRID) + u[1], (float)(sector % GRID) + u[2], v.pos, v.n);
        }
        V3 rgb = g_scene.surf_rgb[h.index];
        V3 brdf = { rgb.x / PI_F, rgb.y / PI_F, rgb.z / PI_F };
        float cos_theta = dot_plain(g_scene.surf[h.index].n, nd);
        cur_brdf = g_scene.surf_lum[h.index] / PI_F;
        thr = { thr.x * ((brdf.x * cos_theta) / pdf), thr.y * ((brdf.y * cos_theta) / pdf), thr.z * ((brdf.z * cos_theta) / pdf) };
        o = { h.pos.x + nd.x * 0.00001f, h.pos.y + nd.y * 0.00001f, h.pos.z + nd.z * 0.00001f };
        d = normalize(nd, c.fma_mode);
    }
    len = c.max_bounces; return { 0.f, 0.f, 0.f };
}

}  // namespace

extern "C" {

void orc_set_max_direction(int on) { g_max_direction = on; }
int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_philox(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t purpose, float* u4) { draw4(seed, pixel, sample, bounce, purpose, u4); }

/* triangles as 9 floats (v0,v1,v2) each; colours 3 floats each */
int orc_scene_set(const float* sv, const float* srgb, int ns, const float* lv, const float* lrgb, int nl) {
    Scene& s = g_scene; s = Scene();
    auto load = [](const float* v, const float* rgb, int n, std::vector<Tri>& tris, std::vector<V3>& cols, std::vector<float>& lums) {
        for (int i = 0; i < n; ++i) {
            const float* p = v + 9 * i; Tri t;
            t.v0 = { p[0], p[1], p[2] }; t.v1 = { p[3], p[4], p[5] }; t.v2 = { p[6], p[7], p[8] }; t.n = tri_normal(t);
            tris.push_back(t); V3 c = { rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2] }; cols.push_back(c); lums.push_back(luminance(c));
        }
    };
    load(sv, srgb, ns, s.surf, s.surf_rgb, s.surf_lum);
    load(lv, lrgb, nl, s.light, s.light_rgb, s.light_lum);
    return 0;
}
int orc_scene_normals(float* snrm, float* slum, float* lnrm, float* llum) {
    for (size_t i = 0; i < g_scene.surf.size(); ++i) { snrm[3 * i] = g_scene.surf[i].n.x; snrm[3 * i + 1] = g_scene.surf[i].n.y; snrm[3 * i + 2] = g_scene.surf[i].n.z; slum[i] = g_scene.surf_lum[i]; }
    for (size_t i = 0; i < g_scene.light.size(); ++i) { lnrm[3 * i] = g_scene.light[i].n.x; lnrm[3 * i + 1] = g_scene.light[i].n.y; lnrm[3 * i + 2] = g_scene.light[i].n.z; llum[i] = g_scene.light_lum[i]; }
    return 0;
}
/* dir is normalised first exactly as Ray::Ray does (ray.cu:6-14) */
int orc_closest_hit(const float* org, const float* dir, int n, int screen_height, int fma_mode, int* type, int* index, float* t, float* pos) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        V3 o = { org[3 * i], org[3 * i + 1], org[3 * i + 2] };
        V3 d = normalize(V3{ dir[3 * i], dir[3 * i + 1], dir[3 * i + 2] }, fma_mode);
        Hit h = closest_hit(o, d, (float)screen_height, fma_mode);
        type[i] = h.type; index[i] = h.index; t[i] = h.t;
        if (pos) { pos[3 * i] = h.pos.x; pos[3 * i + 1] = h.pos.y; pos[3 * i + 2] = h.pos.z; }
    }
    return 0;
}
void orc_map(float x, float y, float* out3) { sc_map(x, y, out3[0], out3[1], out3[2]); }
void orc_grid_dir(const float* gx, const float* gy, int n, const float* pos3, const float* nrm3, float* out) {
    for (int i = 0; i < n; ++i) { V3 d = grid_dir(gx[i], gy[i], V3{ pos3[0], pos3[1], pos3[2] }, V3{ nrm3[0], nrm3[1], nrm3[2] }); out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z; }
}
void orc_uniform_hemisphere(const float* nrm3, float r1, float r2, float* out3) { V3 d = uniform_hemisphere_dir(V3{ nrm3[0], nrm3[1], nrm3[2] }, r1, r2); out3[0] = d.x; out3[1] = d.y; out3[2] = d.z; }
float orc_tri_area(int surface) { return tri_area(g_scene.surf[surface]); }

/* RadianceMap ctor (radiance_map.cu:8-54): counts (:60-67), sampling (:72-84), kd-tree. Returns the volume count. */
int orc_rmap_build(float area_per_sample) {
    srand(1);
    RMap& m = g_rm; m = RMap();
    for (int j = 0; j < (int)g_scene.surf.size(); ++j) {
        int cnt = (int)floor(tri_area(g_scene.surf[j]) / area_per_sample);
        for (int i = 0; i < cnt; ++i) { Volume v; v.pos = sample_on_triangle(g_scene.surf[j]); v.n = g_scene.surf[j].n; v.surface = j; m.vol.push_back(v); }
    }
    int nv = (int)m.vol.size();
    std::vector<int> ids(nv); for (int i = 0; i < nv; ++i) ids[i] = i;
    KdBuild* root = kd_build(ids, 0);
    TreeEl r0; r0.dim = root->dim; r0.leaf = 0; r0.left = r0.right = 0; r0.data = root->median; r0.pos = { 0, 0, 0 }; r0.n = { 0, 0, 0 };
    m.tree.push_back(r0);
    if (nv) kd_flatten(root, m.tree, 0);
    delete root;
    m.q.assign((size_t)nv * A, 0.f); m.cdf.assign((size_t)nv * A, 0.f); m.visits.assign((size_t)nv * A, 0u); m.irr.assign(nv, 0.f);
    m.acc_sum.assign((size_t)nv * A, 0.0); m.acc_cnt.assign((size_t)nv * A, 0u);
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nv; ++i) init_volume_state(i);
    return nv;
}
/* adopt volumes (positions, surface ids) sampled elsewhere -- e.g. downloaded from the product -- and rebuild the tree */
int orc_rmap_counts(int* nvol, int* ntree) { *nvol = (int)g_rm.vol.size(); *ntree = (int)g_rm.tree.size(); return 0; }
int orc_rmap_get_volumes(float* pos, float* nrm, int* surf) {
    for (size_t i = 0; i < g_rm.vol.size(); ++i) {
        const Volume& v = g_rm.vol[i];
        pos[3 * i] = v.pos.x; pos[3 * i + 1] = v.pos.y; pos[3 * i + 2] = v.pos.z; nrm[3 * i] = v.n.x; nrm[3 * i + 1] = v.n.y; nrm[3 * i + 2] = v.n.z; surf[i] = v.surface;
    }
    return 0;
}
int orc_rmap_get_tree(int* dim, int* leaf, unsigned* left, unsigned* right, float* data, float* pos, float* nrm) {
    for (size_t i = 0; i < g_rm.tree.size(); ++i) {
        const TreeEl& e = g_rm.tree[i];
        dim[i] = e.dim; leaf[i] = e.leaf; left[i] = e.left; right[i] = e.right; data[i] = e.data;
        pos[3 * i] = e.pos.x; pos[3 * i + 1] = e.pos.y; pos[3 * i + 2] = e.pos.z; nrm[3 * i] = e.n.x; nrm[3 * i + 1] = e.n.y; nrm[3 * i + 2] = e.n.z;
    }
    return 0;
}
int orc_rmap_get_state(float* q, float* cdf, unsigned* visits, float* irr) {
    if (q) memcpy(q, g_rm.q.data(), g_rm.q.size() * 4);
    if (cdf) memcpy(cdf, g_rm.cdf.data(), g_rm.cdf.size() * 4);
    if (visits) memcpy(visits, g_rm.visits.data(), g_rm.visits.size() * 4);
    if (irr) memcpy(irr, g_rm.irr.data(), g_rm.irr.size() * 4);
    return 0;
}
int orc_rmap_get_acc(double* sum, unsigned* cnt) {
    if (sum) memcpy(sum, g_rm.acc_sum.data(), g_rm.acc_sum.size() * 8);
    if (cnt) memcpy(cnt, g_rm.acc_cnt.data(), g_rm.acc_cnt.size() * 4);
    return 0;
}
int orc_rmap_set_acc(const double* sum, const unsigned* cnt) {
    for (size_t i = 0; i < g_rm.acc_sum.size(); ++i) { g_rm.acc_sum[i] = sum[i]; g_rm.acc_cnt[i] = cnt[i]; }
    return 0;
}
int orc_rmap_set_q(const float* q) { memcpy(g_rm.q.data(), q, g_rm.q.size() * 4); return 0; }
int orc_rmap_update_distributions(void) {
    int nv = (int)g_rm.vol.size();
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nv; ++i) update_distribution(g_rm.vol[i], &g_rm.q[(size_t)i * A], &g_rm.cdf[(size_t)i * A]);
    return 0;
}
int orc_rmap_merge_frame(void) { merge_frame(); return 0; }
int orc_find_closest(const float* pos, const float* nrm, int n, float max_dist, int fma_mode, int* out) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) out[i] = kd_find(V3{ pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] }, V3{ nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2] }, max_dist, fma_mode);
    return 0;
}
int orc_sample_sector(const float* cdf_row, float r, float* pdf) { return sample_sector(cdf_row, r, *pdf); }
int orc_sample_sector_clamped(const float* cdf_row, float r, float* pdf) { int s = sample_sector(cdf_row, r, *pdf); return s >= 0 ? s : sample_sector_fallback(cdf_row, r, *pdf); }
void orc_philox_raw(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { uint32_t c[4] = { ctr4[0], ctr4[1], ctr4[2], ctr4[3] }; philox4x32_10(key2[0], key2[1], c); for (int i = 0; i < 4; ++i) out4[i] = c[i]; }
void orc_cell_cos(int vol, float* out144) { for (int k = 0; k < A; ++k) out144[k] = cell_cos(g_rm.vol[vol], k / GRID, k % GRID); }

/* cfg: see struct Cfg. One frame = spp samples for every pixel, samples numbered sample0 .. sample0+spp-1.
 * out_rgb: 3*W*H floats, x-major (pixel = x*H + y), SUM over the frame's samples (caller divides).
 * stats: [total_path_length, zero_contribution_paths, failed_samples, paths]. */
int orc_render_frame(int method, int width, int height, int spp, int sample0, int max_bounces, float env, unsigned seed,
                     const float* cam3, float yaw_y, float yaw_x, int fma_mode, int td_mode, int clamp_last_bin, float max_dist,
                     float* out_rgb, double* stats) {
    Cfg c; c.width = width; c.height = height; c.spp = spp; c.max_bounces = max_bounces; c.env = env; c.seed = seed;
    c.cam[0] = cam3[0]; c.cam[1] = cam3[1]; c.cam[2] = cam3[2]; c.yaw_y = yaw_y; c.yaw_x = yaw_x;
    c.fma_mode = fma_mode; c.td_mode = td_mode; c.clamp_last_bin = clamp_last_bin; c.max_dist = max_dist;
    long long total_len = 0, zero = 0, nfail = 0;
    #pragma omp parallel for schedule(dynamic, 16) reduction(+ : total_len, zero, nfail)
    for (int pix = 0; pix < width * height; ++pix) {
        int px = pix / height, py = pix % height;
        double acc[3] = { 0, 0, 0 };
        for (int s = 0; s < spp; ++s) {
            int len = 0; bool failed = false; V3 L;
            if (method == 0) L = trace_default(c, (uint32_t)pix, px, py, (uint32_t)(sample0 + s), len);
            else L = trace_sarsa(c, (uint32_t)pix, px, py, (uint32_t)(sample0 + s), len, failed);
            total_len += len; if (failed) nfail++;
            if ((L.x + L.y + L.z) / 3.f < 0.0001f) zero++;         /* THROUGHPUT_THRESHOLD, reinforcement_path_tracing.cu:39-42 */
            if (std::isfinite(L.x) && std::isfinite(L.y) && std::isfinite(L.z)) { acc[0] += L.x; acc[1] += L.y; acc[2] += L.z; }
        }
        out_rgb[3 * pix] = (float)acc[0]; out_rgb[3 * pix + 1] = (float)acc[1]; out_rgb[3 * pix + 2] = (float)acc[2];
    }
    if (stats) { stats[0] = (double)total_len; stats[1] = (double)zero; stats[2] = (double)nfail; stats[3] = (double)width * height * spp; }
    return 0;
}

}  /* extern "C" */
