// rlpt_kernels.cu -- the hot-path kernels (sm_100a): closest hit, wavefront path tracing with ray compaction,
// Expected-SARSA sampling / TD accumulation, Q merge + CDF rebuild, pixel accumulation.
//
// Design (DESIGN.md has the long form):
//  * per bounce two kernels: k_isect (closest hit only, 40 registers, issue-bound on the FP32 pipes) and k_shade (nearest
//    radiance volume, TD target, direction sampling, compaction: latency-bound, no triangle data); the scene (SoA float4
//    triangles, parallelogram scan units, BVH nodes) is staged in shared memory by every k_isect CTA with vectorised float4
//    loads when it fits (otherwise everything comes through the read-only path and L1); B200 has no RT cores, traversal is
//    FP32-pipe work
//  * scenes above 64 primitives go through the BVH (k_isect_bvh): lanes take a new ray from the sub-queue as theirs finishes,
//    and leaf triangles are solved for the whole warp from a shared-memory list, merged by a 64-bit atomicMin on (t, primitive)
//  * one resident wave per kernel; the live-path count of each bounce lives in device memory (counts[bounce][sub-queue]),
//    so a frame is a launch sequence with no host round trip; once few paths are left one run-to-completion launch
//    (k_bounce<TAIL>) finishes them
//  * live rays are compacted every bounce with ballot/popc + one atomicAdd per warp into 32 independent sub-queues
//  * TD targets are accumulated as (sum, count) per (volume, sector) with __match_any_sync warp aggregation
//  * radiance is accumulated into a float4-per-pixel buffer in HBM with one vector RED per terminated path
#include "rlpt_internal.h"

namespace rlpt {

__constant__ float c_cell_cos[CELLS];
void upload_cell_cos(const float* cos144) { cudaMemcpyToSymbol(c_cell_cos, cos144, sizeof(float) * CELLS); }

// ------------------------------------------------------------------------------------------------ scene access
extern __shared__ float4 s_scene[];

template <bool STAGED>
struct SceneView {
    const float4* tri_s; const float4* shade_s; const float4* nodes_s; const float4* scan_s; const int* gid_s;
    const float4* tri_g; const float4* shade_g; const float4* nodes_g;
    int n_tri, n_surf, brute, det_small, n_units, n_items;
    float k1, k2, k3, vmax;
    // brute force: record of primitive `i`; BVH: record at leaf-order position `i` (tri4, primitive id in the third float4's z)
    __device__ __forceinline__ float4 tri(int i) const { return STAGED ? tri_s[i] : __ldg(tri_g + i); }
    __device__ __forceinline__ float4 shade(int i) const { return STAGED ? shade_s[i] : __ldg(shade_g + i); }
    __device__ __forceinline__ float4 node(int i) const {
        return STAGED ? nodes_s[i] : __ldg(nodes_g + i);       // a tree that does not fit is served by L1: staging its top costs a compare-and-select per load (-5 %)
    }
};

// Every CTA copies the scene into shared memory once when it fits (vectorised 16-byte loads, coalesced): brute-force scenes
// their triangles, scan units and slot table, BVH scenes the leaf-ordered triangles and the 4-wide nodes.
template <bool STAGED, bool SHADE = true>
__device__ __forceinline__ SceneView<STAGED> stage_scene(const SceneDev& sc) {
    SceneView<STAGED> v;
    v.tri_g = sc.brute ? sc.tri : sc.tri4; v.shade_g = sc.shade; v.nodes_g = sc.bvh4; v.n_tri = sc.n_tri; v.n_surf = sc.n_surf; v.brute = sc.brute; v.det_small = sc.det_small;
    v.n_units = sc.n_units; v.n_items = sc.n_items; v.k1 = sc.k1; v.k2 = sc.k2; v.k3 = sc.k3; v.vmax = sc.vmax;
    v.tri_s = v.shade_s = v.nodes_s = v.scan_s = nullptr; v.gid_s = nullptr;
    if (STAGED) {
        float4* p = s_scene;
        const int nt = 3 * sc.n_tri, ns = SHADE ? 4 * sc.n_tri : 0, nn = sc.brute ? 0 : 7 * sc.n_nodes4;
        v.tri_s = p; v.shade_s = p + nt; v.nodes_s = p + nt + ns;
        for (int i = threadIdx.x; i < nt; i += blockDim.x) p[i] = __ldg(v.tri_g + i);
        for (int i = threadIdx.x; i < ns; i += blockDim.x) p[nt + i] = __ldg(sc.shade + i);
        for (int i = threadIdx.x; i < nn; i += blockDim.x) p[nt + ns + i] = __ldg(sc.bvh4 + i);
        // brute-force scenes: the parallelogram units of the conservative pre-test (4 float4 each), then the slot -> primitive table
        const int nu = sc.brute ? 4 * sc.n_items : 0, ng = (sc.brute && sc.n_units > 0) ? (sc.n_tri + 3) / 4 : 0;
        v.scan_s = p + nt + ns + nn; v.gid_s = reinterpret_cast<const int*>(p + nt + ns + nn + nu);
        for (int i = threadIdx.x; i < nu; i += blockDim.x) p[nt + ns + nn + i] = __ldg(sc.scan + i);
        for (int i = threadIdx.x; i < ng; i += blockDim.x) p[nt + ns + nn + nu + i] = __ldg(reinterpret_cast<const float4*>(sc.slot_gid) + i);
        __syncthreads();
    }
    return v;
}

__device__ __forceinline__ size_t scene_smem_bytes_dev(const SceneDev& sc) {
    return sizeof(float4) * ((size_t)3 * sc.n_tri + (size_t)4 * sc.n_tri + (sc.brute ? 0 : (size_t)7 * sc.n_nodes4) + (sc.brute ? (size_t)4 * sc.n_items : 0) +
                             ((sc.brute && sc.n_units > 0) ? (size_t)(sc.n_tri + 3) / 4 : 0));
}
// dynamic shared memory of the staged scene (the kernels that do not stage the shading records simply leave that part unused)
size_t scene_smem_bytes(const SceneDev& sc) {
    if (!sc.staged) return 0;
    return sizeof(float4) * ((size_t)3 * sc.n_tri + (size_t)4 * sc.n_tri + (sc.brute ? 0 : (size_t)7 * sc.n_nodes4) + (sc.brute ? (size_t)4 * sc.n_items : 0) +
                             ((sc.brute && sc.n_units > 0) ? (size_t)(sc.n_tri + 3) / 4 : 0));
}

// ------------------------------------------------------------------------------------------------ closest hit
template <bool STAGED>
__device__ __forceinline__ TriRec load_tri(const SceneView<STAGED>& v, int gid) {
    float4 a = v.tri(3 * gid), b = v.tri(3 * gid + 1), c = v.tri(3 * gid + 2);
    return TriRec{ a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y };
}

// ---- 4-wide BVH (rlpt_bvh.cu, k_collapse4). Per-ray constants of the slab test: inv = 1 / (dir * SCREEN_HEIGHT), noi = -(o * inv)
// and the ray's octant as record offsets (near plane = lo where the direction is non-negative). The test is the FMA form
// t = plane * inv - o * inv: against (plane - o) * inv it moves a plane by at most |o| 2^-24 (the rounding of o * inv, seen from the
// plane's side) -- four orders of magnitude inside the padding every leaf box carries (1e-4 + 1/128 of the primitive's extent),
// so a primitive the exact solve accepts is still never culled. A direction component of exactly 0 gives inv = inf and NaN where
// plane and origin have the same sign; fminf / fmaxf drop NaN operands, i.e. that axis is then ignored: conservative as well.
struct Ray4 { float ix, iy, iz, nox, noy, noz; unsigned sx, sy, sz; };        // sx, sy, sz: record offset of the NEAR plane (0 = lo, 1 = hi)
__device__ __forceinline__ Ray4 ray4_setup(float ox, float oy, float oz, float sdx, float sdy, float sdz) {
    Ray4 r; r.ix = 1.f / sdx; r.iy = 1.f / sdy; r.iz = 1.f / sdz;
    r.nox = -(ox * r.ix); r.noy = -(oy * r.iy); r.noz = -(oz * r.iz);
    r.sx = sdx < 0.f ? 1u : 0u; r.sy = sdy < 0.f ? 1u : 0u; r.sz = sdz < 0.f ? 1u : 0u;
    return r;
}
// the four children of node `cur`: entry distances (clamped at 0) and hit flags; links returned as loaded
template <bool STAGED>
__device__ __forceinline__ void bvh4_boxes(const SceneView<STAGED>& v, int cur, const Ray4& r, float best_t, float (&tn)[4], bool (&hit)[4], int (&lk)[4]) {
    const float4* __restrict__ nb = (STAGED ? v.nodes_s : v.nodes_g) + 7u * (unsigned)cur;          // one 64-bit address; the records are 32-bit offsets from it
    auto rec = [&](unsigned i) { return STAGED ? nb[i] : __ldg(nb + i); };
    const float4 nx = rec(r.sx), fx = rec(r.sx ^ 1u), ny = rec(2u + r.sy), fy = rec(2u + (r.sy ^ 1u));
    const float4 nz = rec(4u + r.sz), fz = rec(4u + (r.sz ^ 1u)), l = rec(6u);
    const float nxx[4] = { nx.x, nx.y, nx.z, nx.w }, fxx[4] = { fx.x, fx.y, fx.z, fx.w }, nyy[4] = { ny.x, ny.y, ny.z, ny.w }, fyy[4] = { fy.x, fy.y, fy.z, fy.w };
    const float nzz[4] = { nz.x, nz.y, nz.z, nz.w }, fzz[4] = { fz.x, fz.y, fz.z, fz.w };
    lk[0] = __float_as_int(l.x); lk[1] = __float_as_int(l.y); lk[2] = __float_as_int(l.z); lk[3] = __float_as_int(l.w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = fmaxf(fmaxf(fmaf(nxx[k], r.ix, r.nox), fmaf(nyy[k], r.iy, r.noy)), fmaxf(fmaf(nzz[k], r.iz, r.noz), 0.f));
        const float f = fminf(fminf(fmaf(fxx[k], r.ix, r.nox), fmaf(fyy[k], r.iy, r.noy)), fminf(fmaf(fzz[k], r.iz, r.noz), best_t));
        tn[k] = a; hit[k] = a <= f;
    }
}
// children of a node: the first two slots are always used, unused slots come last (k_collapse4)
__device__ __forceinline__ unsigned bvh4_children(const int (&lk)[4]) { return 2u + (lk[2] != BVH4_EMPTY ? 1u : 0u) + (lk[3] != BVH4_EMPTY ? 1u : 0u); }
// Inner children that were hit, ordered by entry distance. Key = entry distance with the slot number in its two lowest mantissa
// bits (entry distances are >= 0, so the bit patterns order like the values; rounding the distance down by two bits only makes a
// later cull more conservative); 0xffffffff = no child. Five compare-exchanges on integer min / max.
__device__ __forceinline__ void bvh4_sort(unsigned (&key)[4]) {
#define RLPT_CE(i, j) { const unsigned lo_ = min(key[i], key[j]), hi_ = max(key[i], key[j]); key[i] = lo_; key[j] = hi_; }
    RLPT_CE(0, 1) RLPT_CE(2, 3) RLPT_CE(0, 2) RLPT_CE(1, 3) RLPT_CE(1, 2)
#undef RLPT_CE
}
constexpr unsigned B4_NONE = 0xffffffffu;

// Per-lane traversal of the 4-wide tree with the exact solve done where a leaf is found (stack in local memory): the generic
// closest_hit used by the run-to-completion kernel, the Neural-Q tracers and the debug view. The wavefront's own closest-hit
// kernel (k_isect_bvh) and the parity entry point walk the same tree warp-cooperatively (bvh4_trace_warp, below).
template <bool STAGED, bool COUNT>
__device__ __forceinline__ void bvh4_closest_hit(const SceneView<STAGED>& v, float ox, float oy, float oz, float a0, float a1, float a2, float sdx, float sdy, float sdz,
                                                 float& best_t, int& best_gid, unsigned& n_tri, unsigned& n_box) {
    const Ray4 r = ray4_setup(ox, oy, oz, sdx, sdy, sdz);
    uint2 stack[96]; int top = 0; int cur = 0;
    while (true) {
        float tn[4]; bool hit[4]; int lk[4];
        bvh4_boxes<STAGED>(v, cur, r, best_t, tn, hit, lk);
        if (COUNT) n_box += bvh4_children(lk);
        unsigned key[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            key[k] = (hit[k] && lk[k] >= 0) ? ((__float_as_uint(tn[k]) & ~3u) | (unsigned)k) : B4_NONE;
            if (hit[k] && lk[k] < 0) {
                const int first = (~lk[k]) >> 3, cnt = (~lk[k]) & 7;
                for (int j = 0; j < cnt; ++j) {
                    const float4 q0 = v.tri(3 * (first + j)), q1 = v.tri(3 * (first + j) + 1), q2 = v.tri(3 * (first + j) + 2);
                    const TriRec tr{ q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y }; const int gid = __float_as_int(q2.z); float t;
                    if (COUNT) n_tri++;
                    if (tri_solve(tr, ox, oy, oz, a0, a1, a2, best_t, t) && (t < best_t || (t == best_t && gid < best_gid))) { best_t = t; best_gid = gid; }
                }
            }
        }
        bvh4_sort(key);
        if (key[0] != B4_NONE) {
            if (key[3] != B4_NONE && top < 96) stack[top++] = make_uint2((unsigned)lk[0] + (key[3] & 3u), key[3]);
            if (key[2] != B4_NONE && top < 96) stack[top++] = make_uint2((unsigned)lk[0] + (key[2] & 3u), key[2]);
            if (key[1] != B4_NONE && top < 96) stack[top++] = make_uint2((unsigned)lk[0] + (key[1] & 3u), key[1]);
            cur = lk[0] + (int)(key[0] & 3u);
        } else {
            bool found = false;
            while (top > 0) { const uint2 e = stack[--top]; if (__uint_as_float(e.y & ~3u) <= best_t) { cur = (int)e.x; found = true; break; } }
            if (!found) return;
        }
    }
}

// Phase 1 of the brute-force scan over the scan units (parallelogram pairs, then single triangles): the slots no early out can
// reject, as a bit mask in SLOT order (slot -> primitive through gid_s). Branch-free over the whole scene.
template <bool STAGED>
__device__ __forceinline__ unsigned long long unit_scan_mask(const SceneView<STAGED>& v, float ox, float oy, float oz, float a0, float a1, float a2) {
    const float A_ = fmaxf(fabsf(a0), fmaxf(fabsf(a1), fabsf(a2)));
    const float Bm = fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))) + v.vmax;
    const float del = A_ * (v.k1 * Bm + v.k2), kx = (3.f * del) / (v.k3 * Bm);
    unsigned long long mask = 0ull;
    int u = 0;
    for (; u + 2 <= v.n_units; u += 2) {
        unsigned bits = 0u;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4 q0 = v.scan_s[4 * (u + k)], q1 = v.scan_s[4 * (u + k) + 1], q2 = v.scan_s[4 * (u + k) + 2], q3 = v.scan_s[4 * (u + k) + 3];
            const UnitRec r{ q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z };
            bits |= unit_candidates(r, ox, oy, oz, a0, a1, a2, del, kx) << (2 * k);
        }
        mask |= (unsigned long long)bits << (2 * u);
    }
    for (; u < v.n_units; ++u) {
        const float4 q0 = v.scan_s[4 * u], q1 = v.scan_s[4 * u + 1], q2 = v.scan_s[4 * u + 2], q3 = v.scan_s[4 * u + 3];
        const UnitRec r{ q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z };
        mask |= (unsigned long long)unit_candidates(r, ox, oy, oz, a0, a1, a2, del, kx) << (2 * u);
    }
    int slot = 2 * v.n_units;
    if (v.det_small) {
        for (; slot + 4 <= v.n_tri; slot += 4) {
            unsigned bits = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                TriRec r = load_tri(v, v.gid_s[slot + k]);
                if (tri_candidate_small(r, ox, oy, oz, a0, a1, a2)) bits |= 1u << k;
            }
            mask |= (unsigned long long)bits << slot;
        }
        for (; slot < v.n_tri; ++slot) {
            TriRec r = load_tri(v, v.gid_s[slot]);
            if (tri_candidate_small(r, ox, oy, oz, a0, a1, a2)) mask |= 1ull << slot;
        }
    } else {
        for (; slot < v.n_tri; ++slot) {
            TriRec r = load_tri(v, v.gid_s[slot]);
            if (tri_candidate(r, ox, oy, oz, a0, a1, a2)) mask |= 1ull << slot;
        }
    }
    return mask;
}

// Closest hit of one ray: Ray::closest_intersection (G/rays/ray.cu:16-36). (dx,dy,dz) is the normalised direction;
// H = SCREEN_HEIGHT. Result: best_t in the reference's units and the primitive id (-1 = NOTHING). The winner is the
// lexicographic minimum of (t, gid), which is what the reference's scan order with strict < produces.
template <bool STAGED, bool COUNT, bool ALLOW_BVH = true>
__device__ __forceinline__ void closest_hit(const SceneView<STAGED>& v, float ox, float oy, float oz, float dx, float dy, float dz, float H,
                                            float& best_t, int& best_gid, float& sdx, float& sdy, float& sdz, unsigned& n_tri, unsigned& n_box) {
    sdx = RLPT_MUL(dx, H); sdy = RLPT_MUL(dy, H); sdz = RLPT_MUL(dz, H);          // dir * SCREEN_HEIGHT (ray.cu:53)
    const float a0 = RLPT_SUB(0.f, sdx), a1 = RLPT_SUB(0.f, sdy), a2 = RLPT_SUB(0.f, sdz);
    best_t = T_MISS; best_gid = -1;
    if (v.brute) {
        if (v.n_units > 0) {
            // phase 1 marks candidates in SLOT order: slots 2u, 2u+1 = the two triangles of parallelogram u (one conservative
            // pre-test for both, unit_candidates), then one slot per remaining triangle (tri_candidate). phase 2: the exact solve
            // of the marked slots; slots are not in primitive order, so ties go to the lower id explicitly (the reference's
            // scan order with strict <).
            unsigned long long mask = unit_scan_mask<STAGED>(v, ox, oy, oz, a0, a1, a2);
            if (COUNT) n_tri += (unsigned)v.n_tri;
            while (mask) {
                const int sl = __ffsll((long long)mask) - 1; mask &= mask - 1ull;
                const int gid = v.gid_s[sl];
                TriRec r = load_tri(v, gid); float t;
                if (tri_solve(r, ox, oy, oz, a0, a1, a2, best_t, t) && (t < best_t || (t == best_t && gid < best_gid))) { best_t = t; best_gid = gid; }
            }
            return;
        }
        if (v.n_tri <= 64) {
            // two phases: (1) a branch-free pass over every triangle marks the candidates no early out can reject,
            // (2) the full solve with its divisions runs on the marked ones only, in primitive order (so "first lowest id
            // wins a tie" is kept). Lanes disagree about WHICH triangles are candidates, so solving inline would make the
            // whole warp walk the division path at most iterations; here it walks it max-over-lanes(#candidates) times.
            unsigned long long mask = 0ull;
            if (v.det_small) {
                // groups of four: the candidate bits are set at compile-time positions (one predicated LOP3 each) and
                // shifted into the 64-bit mask once per group
                int gid = 0;
                for (; gid + 4 <= v.n_tri; gid += 4) {
                    unsigned bits = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        TriRec r = load_tri(v, gid + k);
                        if (tri_candidate_small(r, ox, oy, oz, a0, a1, a2)) bits |= 1u << k;
                    }
                    mask |= (unsigned long long)bits << gid;
                }
                for (; gid < v.n_tri; ++gid) {
                    TriRec r = load_tri(v, gid);
                    if (tri_candidate_small(r, ox, oy, oz, a0, a1, a2)) mask |= 1ull << gid;
                }
            } else {
#pragma unroll 2
                for (int gid = 0; gid < v.n_tri; ++gid) {
                    TriRec r = load_tri(v, gid);
                    if (tri_candidate(r, ox, oy, oz, a0, a1, a2)) mask |= 1ull << gid;
                }
            }
            if (COUNT) n_tri += (unsigned)v.n_tri;
            while (mask) {
                const int gid = __ffsll((long long)mask) - 1; mask &= mask - 1ull;
                TriRec r = load_tri(v, gid); float t;
                if (tri_solve(r, ox, oy, oz, a0, a1, a2, best_t, t) && t < best_t) { best_t = t; best_gid = gid; }
            }
            return;
        }
        for (int gid = 0; gid < v.n_tri; ++gid) {
            TriRec r = load_tri(v, gid); float t;
            if (COUNT) n_tri++;
            if (tri_solve(r, ox, oy, oz, a0, a1, a2, best_t, t) && t < best_t) { best_t = t; best_gid = gid; }
        }
        return;
    }
    // (k_isect is launched for brute-force scenes only -- BVH scenes go to k_isect_bvh -- and is instantiated without the tree walk: its 96-entry local stack and
    // registers are then not part of a kernel that is tuned to 40 registers)
    if (ALLOW_BVH) bvh4_closest_hit<STAGED, COUNT>(v, ox, oy, oz, a0, a1, a2, sdx, sdy, sdz, best_t, best_gid, n_tri, n_box);
}

// Closest hit for a warp of CAMERA rays: same origin, directions within a pixel or two. Phase 1 is done once per warp instead
// of once per ray: lane j pre-tests scan record j (a parallelogram pair or a single triangle, <= 32 records) against the
// warp's central ray, with the thresholds of unit_candidates widened by how far any lane's direction is from the central one
// -- with a common origin b = o - v0 is shared and Cramer's numerators are LINEAR in a:  D = a . n,  Y = a . (b x e2),
// Z = a . (e1 x b),  so |D(a) - D(ac)| <= sum_i rho_i |n_i| etc. with rho_i = max over lanes |a_i - ac_i|. The ballots of the
// lanes' verdicts are the warp's candidate list, and phase 2 (the exact solve) runs on it in lockstep.
template <bool STAGED>
__device__ __forceinline__ void closest_hit_bundle(const SceneView<STAGED>& v, bool valid, float ox, float oy, float oz, float dx, float dy, float dz, float H,
                                                   float& best_t, int& best_gid, unsigned& n_tri) {
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31;
    const float sdx = RLPT_MUL(dx, H), sdy = RLPT_MUL(dy, H), sdz = RLPT_MUL(dz, H);
    const float a0 = RLPT_SUB(0.f, sdx), a1 = RLPT_SUB(0.f, sdy), a2 = RLPT_SUB(0.f, sdz);
    best_t = T_MISS; best_gid = -1;
    const int src = __ffs(__ballot_sync(full, valid)) - 1;                     // a lane that holds a ray (callers pass whole warps with >= 1 ray)
    const float c0 = __shfl_sync(full, a0, src), c1 = __shfl_sync(full, a1, src), c2 = __shfl_sync(full, a2, src);
    const float r0 = __uint_as_float(__reduce_max_sync(full, __float_as_uint(valid ? fabsf(a0 - c0) : 0.f)));   // non-negative floats order like their bits
    const float r1 = __uint_as_float(__reduce_max_sync(full, __float_as_uint(valid ? fabsf(a1 - c1) : 0.f)));
    const float r2 = __uint_as_float(__reduce_max_sync(full, __float_as_uint(valid ? fabsf(a2 - c2) : 0.f)));
    bool candA = false, candB = false;
    if ((int)lane < v.n_items) {
        const float4 q0 = v.scan_s[4 * lane], q1 = v.scan_s[4 * lane + 1], q2 = v.scan_s[4 * lane + 2], q3 = v.scan_s[4 * lane + 3];
        const float e1x = q0.w, e1y = q1.x, e1z = q1.y, e2x = q1.z, e2y = q1.w, e2z = q2.x, nx = q2.y, ny = q2.z, nz = q2.w;
        const float bx = ox - q0.x, by = oy - q0.y, bz = oz - q0.z;
        const float yx = by * e2z - bz * e2y, yy = bz * e2x - bx * e2z, yz = bx * e2y - by * e2x;        // b x e2
        const float zx = e1y * bz - e1z * by, zy = e1z * bx - e1x * bz, zz = e1x * by - e1y * bx;        // e1 x b
        const float D = c0 * nx + c1 * ny + c2 * nz, X = bx * nx + by * ny + bz * nz;
        const float Y = c0 * yx + c1 * yy + c2 * yz, Z = c0 * zx + c1 * zy + c2 * zz;
        const float eD = r0 * fabsf(nx) + r1 * fabsf(ny) + r2 * fabsf(nz);
        const float eY = r0 * fabsf(yx) + r1 * fabsf(yy) + r2 * fabsf(yz), eZ = r0 * fabsf(zx) + r1 * fabsf(zy) + r2 * fabsf(zz);
        const float A_ = fmaxf(fabsf(c0) + r0, fmaxf(fabsf(c1) + r1, fabsf(c2) + r2));
        const float Bm = fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))) + v.vmax;
        const float del = A_ * (v.k1 * Bm + v.k2), delx = v.k3 * Bm, lim = 3.f * del;
        const float Ds = fabsf(D), Xs = D < 0.f ? -X : X, Ys = D < 0.f ? -Y : Y, Zs = D < 0.f ? -Z : Z, S = Ys + Zs;
        const bool unsure = !(Ds > del + eD), front = Xs >= -delx;
        const bool inA = Ys >= -(lim + eY) && Zs >= -(lim + eZ) && S <= Ds + (lim + eY + eZ + eD);
        const bool inB = Ys <= q3.x * Ds + (lim + eY + eD) && Zs <= q3.y * Ds + (lim + eZ + eD) && S >= q3.z * Ds - (lim + eY + eZ + eD);
        candA = unsure || (front && inA);
        candB = (int)lane < v.n_units && (unsure || (front && inB));
    }
    unsigned mA = __ballot_sync(full, candA), mB = __ballot_sync(full, candB);
    n_tri += (unsigned)v.n_tri;
    while (mA | mB) {
        int gid;
        if (mA) { const int j = __ffs(mA) - 1; mA &= mA - 1u; gid = j < v.n_units ? v.gid_s[2 * j] : v.gid_s[v.n_units + j]; }     // singles: slot 2 n_units + (j - n_units)
        else { const int j = __ffs(mB) - 1; mB &= mB - 1u; gid = v.gid_s[2 * j + 1]; }
        TriRec r = load_tri(v, gid); float t;
        if (valid && tri_solve(r, ox, oy, oz, a0, a1, a2, best_t, t) && (t < best_t || (t == best_t && gid < best_gid))) { best_t = t; best_gid = gid; }
    }
}

template <bool STAGED, bool COUNT>
__global__ void __launch_bounds__(BLOCK) k_closest_hit(SceneDev sc, const float* __restrict__ org, const float* __restrict__ dir, int n, float H,
                                                       int* __restrict__ type, int* __restrict__ index, float* __restrict__ t_out,
                                                       unsigned long long* __restrict__ counters) {
    SceneView<STAGED> v = stage_scene<STAGED, false>(sc);
    unsigned nt = 0, nb = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        f3 d = normalize_ref(f3{ dir[3 * i], dir[3 * i + 1], dir[3 * i + 2] });
        float t, sx, sy, sz; int gid;
        closest_hit<STAGED, COUNT>(v, org[3 * i], org[3 * i + 1], org[3 * i + 2], d.x, d.y, d.z, H, t, gid, sx, sy, sz, nt, nb);
        type[i] = gid < 0 ? 0 : (gid < sc.n_surf ? 2 : 1);
        index[i] = gid < 0 ? -1 : (gid < sc.n_surf ? gid : gid - sc.n_surf);
        t_out[i] = t;
    }
    if (COUNT) {
        nt = __reduce_add_sync(0xffffffffu, nt); nb = __reduce_add_sync(0xffffffffu, nb);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&counters[0], (unsigned long long)nt); atomicAdd(&counters[1], (unsigned long long)nb); }
    }
}

// ------------------------------------------------------------------------------------------------ nearest volume
__device__ __forceinline__ int find_volume(const RadianceDev& rm, float px, float py, float pz, int cls, unsigned& n_kd) {
    const float4* __restrict__ inner = rm.kd_inner; const float4* __restrict__ posn = rm.vol_posn;
    if (rm.vc_table) {
        const int4* __restrict__ tb = rm.vc_table; const float4* __restrict__ cd = rm.vc_cand;
        const float d0 = kd_distance(px, py, pz, rm.root_px, rm.root_py, rm.root_pz);
        int r = vcell_find(rm.vc, [&](uint32_t i, int& c, int& k, int& s, int& n) { int4 e = __ldg(tb + i); c = e.x; k = e.y; s = e.z; n = e.w; },
                           [&](int i, float& x, float& y, float& z, int& v) { float4 p = __ldg(cd + i); x = p.x; y = p.y; z = p.z; v = __float_as_int(p.w); },
                           px, py, pz, cls, d0);
        if (r >= 0) return r;
        if (rm.vx_table) {
            const int4* __restrict__ xt = rm.vx_table; const float4* __restrict__ xc = rm.vx_cand;
            r = vext_find(rm.vc, rm.vx_mask, [&](uint32_t i, int& c, int& k, int& s, int& n) { int4 e = __ldg(xt + i); c = e.x; k = e.y; s = e.z; n = e.w; },
                          [&](int i, float& x, float& y, float& z, int& v, float (&lo)[3], float (&hi)[3]) {
                              float4 a = __ldg(xc + 3 * i), b = __ldg(xc + 3 * i + 1), c = __ldg(xc + 3 * i + 2);
                              x = a.x; y = a.y; z = a.z; v = __float_as_int(a.w); lo[0] = b.x; lo[1] = b.y; lo[2] = b.z; hi[0] = b.w; hi[1] = c.x; hi[2] = c.y;
                          }, px, py, pz, cls, d0, rm.within_abs);
            if (r >= 0) return r;
        }
    }
    // exact reference search: what neither level decides (distance ties, points away from every surface)
    n_kd++;
    return kd_find(
        [&](uint32_t idx, float& split, uint32_t& l, uint32_t& r, int& dim) {
            float4 n = __ldg(inner + idx); split = n.x; l = __float_as_uint(n.y); r = __float_as_uint(n.z); dim = __float_as_int(n.w);
        },
        [&](int vol, float& x, float& y, float& z, int& c) { float4 p = __ldg(posn + vol); x = p.x; y = p.y; z = p.z; c = __float_as_int(p.w); },
        rm.root, rm.root_px, rm.root_py, rm.root_pz, px, py, pz, cls, rm.within_abs);
}

__global__ void __launch_bounds__(BLOCK) k_find_closest(RadianceDev rm, const float* __restrict__ pos, const int* __restrict__ cls, int n, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_kd = 0;
    if (i < n) out[i] = find_volume(rm, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], cls[i], n_kd);
}
void launch_find_closest(const RadianceDev& rm, const SceneDev&, const float* pos, const float* cls_as_float, int n, int* out, cudaStream_t s) {
    k_find_closest<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(rm, pos, reinterpret_cast<const int*>(cls_as_float), n, out);
}

// "zero contribution" of the reference's statistics: mean(rgb) < THROUGHPUT_THRESHOLD = 0.0001 (G/main.cu:223-229). The division by
// three sat in the divergent termination branch (2 % of k_shade's instructions for a counter); fl(x / 3) < 0.0001f holds exactly
// for the floats x <= 0x1.3a92ap-12 (correctly rounded division is monotonic; the boundary was found by stepping through the
// neighbouring floats), so the comparison is made on the sum. NaN compares false in both forms.
__device__ __forceinline__ bool zero_contribution(float lr, float lg, float lb) { return (lr + lg + lb) <= 0x1.3a92ap-12f; }

// ------------------------------------------------------------------------------------------------ warp helpers
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// Warp-aggregated TD accumulation: lanes that update the same (volume, sector) are found with __match_any_sync, their
// targets summed by shuffles, and the group leader issues one RED.ADD.F32 + one RED.ADD.U32.
__device__ __forceinline__ void td_accumulate(const RadianceDev& rm, bool active, uint32_t key, float target) {
    const unsigned full = 0xffffffffu;
    unsigned lane = threadIdx.x & 31;
    uint32_t k = active ? key : (0xffffffffu - lane);             // inactive lanes get unique dummy keys
    unsigned peers = __match_any_sync(full, k);
    if (__all_sync(full, peers == (1u << lane))) {                // common case after compaction: all keys distinct
        if (active) { atomicAdd(rm.acc_sum + key, target); atomicAdd(rm.acc_cnt + key, 1u); }
        return;
    }
    int n = __popc(peers), nmax = __reduce_max_sync(full, n);
    unsigned m = peers; float total = 0.f;
    for (int it = 0; it < nmax; ++it) {
        int src = m ? (__ffs(m) - 1) : (int)lane;
        float val = __shfl_sync(full, target, src);
        if (m) total += val;
        m &= m - 1;
    }
    if (active && lane == (unsigned)(__ffs(peers) - 1)) { atomicAdd(rm.acc_sum + key, total); atomicAdd(rm.acc_cnt + key, (unsigned)n); }
}

// ------------------------------------------------------------------------------------------------ wavefront kernels
// Two pipelines trace a frame (DESIGN.md "Wavefront"):
//   split  per bounce k_isect (closest hit only: ~40 registers, every warp slot of the SM filled, FP32-pipe bound) writes
//          (t, primitive) per queue slot, then k_shade (nearest volume, TD target, direction sampling, compaction:
//          L2-latency bound, no triangle data) consumes it
//   fused  k_bounce does both in one launch; with TAIL it keeps every path in its thread until it terminates -- used once
//          the live-path count is below one wave, where a launch per bounce would only add its latency
#ifndef RLPT_MINBLOCKS
#define RLPT_MINBLOCKS 1
#endif
#ifndef RLPT_CDF_2LEVEL
#define RLPT_CDF_2LEVEL 1
#endif
#ifndef RLPT_SHADE_MINBLOCKS
#define RLPT_SHADE_MINBLOCKS 4
#endif
#ifndef RLPT_ISECT_MINBLOCKS
#define RLPT_ISECT_MINBLOCKS 6
#endif


struct PathState {
    float ox, oy, oz, dx, dy, dz, tr, tg, tb, cur_brdf;
    uint32_t pixel, sample, volsec;
};

// camera sample `i` of this lane: Ray::sample_ray_through_pixel (G/rays/ray.cu:144-172) with Philox counters (pixel, sample)
__device__ __forceinline__ void primary_state(const FrameParams& p, const FrameDyn& dyn, int i, PathState& s) {
    s.pixel = (uint32_t)(i / p.spp); s.sample = dyn.sample_base + (uint32_t)(i % p.spp);
    float u0, u1, u2, u3; draw4(p.seed, s.pixel, s.sample, 0u, PURPOSE_CAMERA, u0, u1, u2, u3);
    f3 d = camera_dir((int)(s.pixel / (uint32_t)p.height), (int)(s.pixel % (uint32_t)p.height), u0, u1, p.width, p.height, dyn.rotated != 0, dyn.cy, dyn.sy, dyn.cx, dyn.sx);
    s.ox = dyn.cam_x; s.oy = dyn.cam_y; s.oz = dyn.cam_z; s.dx = d.x; s.dy = d.y; s.dz = d.z;
    s.tr = s.tg = s.tb = 1.f; s.cur_brdf = 0.f; s.volsec = 0u;
}
// The path queues are write-once / read-once streams of ~0.9 GB per frame; they go through L2 with the evict-first
// policy (ld/st.global.cs) so that they do not push out what is re-read at random: the nearest-volume tables, the CDFs
// and the TD accumulators (~90 MB for Cornell, inside the 126 MB L2).
__device__ __forceinline__ void load_state(const PathQueue& q, int i, PathState& s) {
    float4 a = __ldcs(q.o + i), b = __ldcs(q.d + i), c = __ldcs(q.thr + i); uint32_t m = __ldcs(q.meta + i);
    s.ox = a.x; s.oy = a.y; s.oz = a.z; s.pixel = __float_as_uint(a.w);
    s.dx = b.x; s.dy = b.y; s.dz = b.z; s.cur_brdf = b.w;
    s.tr = c.x; s.tg = c.y; s.tb = c.z; s.volsec = __float_as_uint(c.w); s.sample = m;
}
__device__ __forceinline__ void store_state(const PathQueue& q, int slot, const PathState& s, int bounce) {
    __stcs(q.o + slot, make_float4(s.ox, s.oy, s.oz, __uint_as_float(s.pixel)));
    __stcs(q.d + slot, make_float4(s.dx, s.dy, s.dz, s.cur_brdf));
    __stcs(q.thr + slot, make_float4(s.tr, s.tg, s.tb, __uint_as_float(s.volsec)));
    __stcs(q.meta + slot, s.sample);          // the full 32-bit sample index: it is a Philox counter word, and frames * ranks * spp passes 2^24 within minutes
}
// Sub-queues. A lane's path queue is NSUB independent queues side by side (slots [k * sub_cap, (k + 1) * sub_cap)), each
// with its own live-path counter per bounce and its own family of CTAs (blockIdx % NSUB == k) that both drains it and
// refills its successor. A single queue has a single slot counter, and same-address atomics serialise in L2: with one
// atomicAdd per warp that was a quarter of k_shade's stall samples, with one per CTA the two barriers it needs were a
// fifth. NSUB counters in separate 32-byte sectors take 1/NSUB of the traffic each and need no barrier. Survivors of
// family k never outnumber its inputs, so sub_cap = its share of the primary paths bounds every later bounce.
struct SubQueue { int base; int n; int first; int stride; int* next_count; };
template <bool PRIMARY>
__device__ __forceinline__ SubQueue sub_queue(const FrameParams& p, int bounce) {
    SubQueue q;
    const int k = (int)(blockIdx.x % NSUB), j = (int)(blockIdx.x / NSUB), fam = (int)(gridDim.x / NSUB);
    q.base = k * p.sub_cap;
    if (PRIMARY) { const int total = p.width * p.height * p.spp; q.n = min(p.sub_cap, max(0, total - q.base)); }
    else q.n = p.counts[(bounce * NSUB + k) * COUNT_STRIDE];
    q.first = j * BLOCK + (int)threadIdx.x; q.stride = fam * BLOCK;
    q.next_count = p.counts + ((bounce + 1) * NSUB + k) * COUNT_STRIDE;
    return q;
}
// wavefront compaction: survivors are packed densely into the sub-queue's successor (one atomicAdd per warp)
__device__ __forceinline__ void compact_store(const SubQueue& sq, const PathQueue& qo, int bounce, bool alive, const PathState& s) {
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31;
    const unsigned bal = __ballot_sync(full, alive);
    if (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(sq.next_count, __popc(bal));
        base = __shfl_sync(full, base, 0);
        if (alive) store_state(qo, sq.base + base + __popc(bal & lanemask_lt()), s, bounce + 1);
    }
}

// Everything of one bounce after the closest hit (t, gid) of path `s` is known: TD target of the previous (volume,
// sector), termination into the frame buffer, or the next direction. Returns true when the path goes on; `s` then holds
// the next ray. Called by whole warps (td_accumulate is a warp collective); `valid` marks the lanes that hold a path.
// path_trace_iterative (G/path_tracing/default_path_tracing.cu:36-88) / path_trace_reinforcement_iterative
// (G/path_tracing/reinforcement_path_tracing.cu:48-120).
template <bool SARSA, bool TD_PREV, class ShadeLoad>
__device__ __forceinline__ bool shade_step(const FrameParams& p, const FrameDyn& dyn, const ShadeLoad& shade, int n_surf, int bounce, bool valid,
                                           PathState& s, float t, int gid, unsigned& st_len, unsigned& st_zero, unsigned& st_term, unsigned& n_kd) {
    const bool surface = valid && gid >= 0 && gid < n_surf;
    const bool light = valid && gid >= n_surf;
    float hx = 0, hy = 0, hz = 0; float4 sN = make_float4(0, 1, 0, 0), sT = make_float4(1, 0, 0, 0);
    int nv = 0;
    if (surface) {
        const float H = (float)p.height;
        const float sdx = RLPT_MUL(s.dx, H), sdy = RLPT_MUL(s.dy, H), sdz = RLPT_MUL(s.dz, H);      // dir * SCREEN_HEIGHT (ray.cu:53)
        hx = RLPT_FMA(sdx, t, s.ox); hy = RLPT_FMA(sdy, t, s.oy); hz = RLPT_FMA(sdz, t, s.oz);      // position = start + t*dir (ray.cu:65)
        sN = shade(4 * gid); sT = shade(4 * gid + 1);
        if (SARSA) nv = find_volume(p.rm, hx, hy, hz, __float_as_int(sT.w), n_kd);
    }
    if (SARSA && TD_PREV) {
        // RadianceMap::temporal_difference_update_radiance_volume_sector (radiance_map.cu:111-146): target by hit type
        float target = 0.f;
        if (surface) target = __ldg(p.rm.irradiance + nv) * ((2.f * PI_F) / 144.f) * s.cur_brdf;   // get_irradiance_estimate (radiance_volume.cu:305-307)
        else if (light) target = s.cur_brdf * shade(4 * gid).w;
        else target = s.cur_brdf * p.env;
        td_accumulate(p.rm, valid && dyn.learn, (s.volsec >> 8) * (uint32_t)CELLS + (s.volsec & 0xffu), target);
    }
    if (valid && !surface) {
        // NOTHING: throughput * ENVIRONMENT_LIGHT; AREA_LIGHT: throughput * diffuse_p (default_path_tracing.cu:52-63)
        float lr = p.env, lg = p.env, lb = p.env;
        if (light) { float4 e = shade(4 * gid + 3); lr = e.x; lg = e.y; lb = e.z; }
        lr *= s.tr; lg *= s.tg; lb *= s.tb;
        if (lr != 0.f || lg != 0.f || lb != 0.f) atomicAdd(p.accum + s.pixel, make_float4(lr, lg, lb, 0.f));
        st_len += (unsigned)bounce + 1u; st_term++;
        if (zero_contribution(lr, lg, lb)) st_zero++;                                         // THROUGHPUT_THRESHOLD
        return false;
    }
    if (!surface) return false;
    if (bounce + 1 >= p.max_bounces) {            // the reference's loop ends here and returns vec3(0), path_length = MAX_RAY_BOUNCES
        st_len += (unsigned)p.max_bounces; st_term++; st_zero++;
        return false;
    }
    float4 sB = shade(4 * gid + 2), sC = shade(4 * gid + 3);
    f3 N = { sN.x, sN.y, sN.z }, T = { sT.x, sT.y, sT.z }, B = { sB.x, sB.y, sB.z };
    float u0, u1, u2, u3; draw4(p.seed, s.pixel, s.sample, (uint32_t)bounce, PURPOSE_BOUNCE, u0, u1, u2, u3);
    f3 nd; float scale;
    if (SARSA) {
        // importance_sample_ray_direction -> sample_direction_from_radiance_distribution (radiance_volume.cu:192-244)
        const float* __restrict__ row = p.rm.cdf + (size_t)nv * CELLS;
        float pdf; int sector;
        if (dyn.max_dir) {
            // RadianceVolume::sample_max_direction_from_radiance_distribution (radiance_volume.cu:248-278): the first cell holding the largest Q,
            // pdf from the width of its CDF bin (cell 0: its proper width cdf[0]; the reference subtracts cdf[0] from itself there)
            const float* __restrict__ qrow = p.rm.q + (size_t)nv * CELLS;
            sector = 0; float mx = __ldg(qrow);
            for (int k = 1; k < CELLS; ++k) { const float qk = __ldg(qrow + k); if (mx < qk) { mx = qk; sector = k; } }
            pdf = RHO * ((__ldg(row + sector) - (sector ? __ldg(row + sector - 1) : 0.f)) / GRID_RHO);
        } else {
#if RLPT_CDF_2LEVEL
        const float4* __restrict__ rows4 = reinterpret_cast<const float4*>(p.rm.cdf_rows + (size_t)nv * GRID);
        const float4* __restrict__ row4 = reinterpret_cast<const float4*>(row);
        sector = sample_sector_2level(
            [&](float (&e)[12]) { float4 a = __ldg(rows4), b = __ldg(rows4 + 1), c = __ldg(rows4 + 2); e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w; e[8] = c.x; e[9] = c.y; e[10] = c.z; e[11] = c.w; },
            [&](int j, float (&e)[12]) { float4 a = __ldg(row4 + 3 * j), b = __ldg(row4 + 3 * j + 1), c = __ldg(row4 + 3 * j + 2); e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w; e[8] = c.x; e[9] = c.y; e[10] = c.z; e[11] = c.w; },
            [&](int k) { return __ldg(row + k); }, u0, pdf);
#else
        sector = sample_sector([&](int k) { return __ldg(row + k); }, u0, pdf);
#endif
        }
        nd = grid_to_direction((float)(sector / GRID) + u1, (float)(sector % GRID) + u2, T, N, B);
        float cos_theta = N.x * nd.x + N.y * nd.y + N.z * nd.z;                       // reinforcement_path_tracing.cu:107
        scale = cos_theta / pdf;
        s.volsec = ((uint32_t)nv << 8) | (uint32_t)sector;
        s.cur_brdf = sN.w;                                                            // material.luminance / pi (:109)
    } else {
        nd = uniform_hemisphere(u0, u1, T, N, B);                                     // cos_theta = u0, pdf = RHO
        scale = u0 / RHO;
    }
    s.tr *= sC.x * scale; s.tg *= sC.y * scale; s.tb *= sC.z * scale;
    s.ox = RLPT_FMA(RAY_EPS, nd.x, hx); s.oy = RLPT_FMA(RAY_EPS, nd.y, hy); s.oz = RLPT_FMA(RAY_EPS, nd.z, hz);
    f3 nn = normalize_ref(nd); s.dx = nn.x; s.dy = nn.y; s.dz = nn.z;
    return true;
}

__device__ __forceinline__ void flush_path_stats(const FrameParams& p, unsigned st_len, unsigned st_zero, unsigned st_term, unsigned n_kd) {
    const unsigned full = 0xffffffffu;
    st_len = __reduce_add_sync(full, st_len); st_zero = __reduce_add_sync(full, st_zero); st_term = __reduce_add_sync(full, st_term); n_kd = __reduce_add_sync(full, n_kd);
    if ((threadIdx.x & 31) == 0 && n_kd) atomicAdd(p.stats + 5, (unsigned long long)n_kd);
    if ((threadIdx.x & 31) == 0 && st_term) {
        atomicAdd(p.stats + 0, (unsigned long long)st_len); atomicAdd(p.stats + 1, (unsigned long long)st_zero); atomicAdd(p.stats + 2, (unsigned long long)st_term);
    }
}
// work counters for the roofline: triangle tests and box tests actually executed; two integer adds per
// test next to ~72 / ~18 FP32 operations, so they stay on in the product build
__device__ __forceinline__ void flush_work_counters(const FrameParams& p, unsigned n_tri, unsigned n_box) {
    const unsigned full = 0xffffffffu;
    n_tri = __reduce_add_sync(full, n_tri); n_box = __reduce_add_sync(full, n_box);
    if ((threadIdx.x & 31) == 0 && (n_tri | n_box)) { atomicAdd(p.stats + 3, (unsigned long long)n_tri); atomicAdd(p.stats + 4, (unsigned long long)n_box); }
}
__device__ __forceinline__ void capture_ray(const FrameParams& p, const FrameDyn& dyn, int bounce, bool valid, const PathState& s) {
    if (dyn.capture_max > 0 && bounce == dyn.capture_bounce && valid) {
        int slot = atomicAdd(p.capture_n, 1);
        if (slot < dyn.capture_max) { p.capture_o[slot] = make_float4(s.ox, s.oy, s.oz, 0.f); p.capture_d[slot] = make_float4(s.dx, s.dy, s.dz, 0.f); }
    }
}

// Fused: one bounce of one path per thread (PRIMARY: the path is generated here), or with TAIL every remaining bounce.
template <bool STAGED, bool SARSA, bool PRIMARY, bool TAIL>
__global__ void __launch_bounds__(BLOCK, RLPT_MINBLOCKS) k_bounce(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, int bounce) {
    const SubQueue sq = sub_queue<PRIMARY>(p, bounce);
    const int n_round = (sq.n + 31) & ~31;                        // whole warps stay together for the collectives
    if ((int)(blockIdx.x / NSUB) * BLOCK >= n_round) return;      // nothing for this CTA: do not even stage the scene
    SceneView<STAGED> v = stage_scene<STAGED, true>(p.scene);
    const unsigned full = 0xffffffffu;
    const PathQueue qi = p.q[bounce & 1], qo = p.q[(bounce + 1) & 1];
    unsigned st_len = 0, st_zero = 0, st_term = 0, n_tri = 0, n_box = 0, n_kd = 0;
    const float H = (float)p.height;
    auto shade = [&](int k) { return v.shade(k); };
    for (int i = sq.first; i < n_round; i += sq.stride) {
        bool live = i < sq.n;
        PathState s{};
        if (live) {
            if (PRIMARY) {
                primary_state(p, dyn, sq.base + i, s);
                if ((sq.base + i) % p.spp == 0) atomicAdd(&p.accum[s.pixel].w, (float)p.spp);          // samples accumulated for this pixel
            } else load_state(qi, sq.base + i, s);
        }
        int b = bounce;
        while (true) {
            capture_ray(p, dyn, b, live, s);
            float t = T_MISS, sdx, sdy, sdz; int gid = -1;
            if (live) closest_hit<STAGED, true>(v, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, H, t, gid, sdx, sdy, sdz, n_tri, n_box);
            if (PRIMARY) live = shade_step<SARSA, false>(p, dyn, shade, v.n_surf, b, live, s, t, gid, st_len, st_zero, st_term, n_kd);
            else live = shade_step<SARSA, true>(p, dyn, shade, v.n_surf, b, live, s, t, gid, st_len, st_zero, st_term, n_kd);
            if (!TAIL) break;
            ++b;
            if (!__any_sync(full, live)) break;
        }
        if (!TAIL) compact_store(sq, qo, bounce, live, s);
    }
    flush_path_stats(p, st_len, st_zero, st_term, n_kd);
    flush_work_counters(p, n_tri, n_box);
}

// (t, primitive) as one 64-bit key whose unsigned order is the lexicographic order of (|t|, primitive id): results of exact solves
// done by any lane in any order merge with atomicMin, and the minimum is what the reference's scan order with strict < keeps.
constexpr unsigned long long KEY_MISS = ((unsigned long long)0x497423F0u << 32) | 0xffffffffull;    // bits(999999.f), primitive -1
static_assert(BLOCK % 32 == 0, "whole warps");
__device__ __forceinline__ unsigned long long hit_key(float t, int gid) {
    const unsigned tb = __float_as_uint(t);
    return ((unsigned long long)(tb & 0x7fffffffu) << 32) | (unsigned long long)(((unsigned)gid << 1) | (tb >> 31));
}

// Split, first half: the closest hit of every ray of queue `bounce` -> hit[slot] = (t, as_float(primitive id))
template <bool STAGED, bool PRIMARY>
__global__ void __launch_bounds__(BLOCK, RLPT_ISECT_MINBLOCKS) k_isect(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, int bounce) {
    const SubQueue sq = sub_queue<PRIMARY>(p, bounce);
    if ((int)(blockIdx.x / NSUB) * BLOCK >= sq.n) return;
    SceneView<STAGED> v = stage_scene<STAGED, false>(p.scene);
    const PathQueue qi = p.q[bounce & 1];
    unsigned n_tri = 0, n_box = 0;
    const float H = (float)p.height;
    if (PRIMARY && v.brute && p.scene.bundle && v.n_items > 0 && v.n_items <= 32) {
        // camera rays: one pre-test per warp (closest_hit_bundle); whole warps stay together for its collectives
        const int n_round = (sq.n + 31) & ~31;
        for (int i = sq.first; i < n_round; i += sq.stride) {
            const bool valid = i < sq.n;
            PathState s{}; if (valid) primary_state(p, dyn, sq.base + i, s);
            const float cx = dyn.cam_x, cy = dyn.cam_y, cz = dyn.cam_z;            // the shared origin (primary_state sets exactly this)
            if (valid && dyn.capture_max > 0 && bounce == dyn.capture_bounce) {
                int slot = atomicAdd(p.capture_n, 1);
                if (slot < dyn.capture_max) { p.capture_o[slot] = make_float4(cx, cy, cz, 0.f); p.capture_d[slot] = make_float4(s.dx, s.dy, s.dz, 0.f); }
            }
            float t; int gid;
            closest_hit_bundle<STAGED>(v, valid, cx, cy, cz, s.dx, s.dy, s.dz, H, t, gid, n_tri);
            if (valid) __stcs(p.hit + sq.base + i, make_float2(t, __int_as_float(gid)));
        }
        flush_work_counters(p, n_tri, n_box);
        return;
    }
    // (Tried and dropped: the exact solves of the scan's candidates for the whole warp at once -- candidates of all 32 rays in a
    // shared-memory list placed by a warp prefix sum, solved 32 at a time with the owner's ray fetched by shuffles and merged by
    // the 64-bit atomicMin k_isect_bvh uses. Bit-exact, lanes per instruction in phase 2 go from 15 to ~30, but the list
    // bookkeeping, seven shuffles per entry, the compare-and-swap loop behind a 64-bit shared-memory atomicMin and the lost
    // "farther than the current best" early-out cost more than the idle lanes did: 2175 -> 2097 Mpaths/s.)
    for (int i = sq.first; i < sq.n; i += sq.stride) {
        float ox, oy, oz, dx, dy, dz;
        if (PRIMARY) {
            PathState s; primary_state(p, dyn, sq.base + i, s);
            ox = s.ox; oy = s.oy; oz = s.oz; dx = s.dx; dy = s.dy; dz = s.dz;
        } else {
            float4 a = __ldcs(qi.o + sq.base + i), b = __ldcs(qi.d + sq.base + i);
            ox = a.x; oy = a.y; oz = a.z; dx = b.x; dy = b.y; dz = b.z;
        }
        if (dyn.capture_max > 0 && bounce == dyn.capture_bounce) {
            int slot = atomicAdd(p.capture_n, 1);
            if (slot < dyn.capture_max) { p.capture_o[slot] = make_float4(ox, oy, oz, 0.f); p.capture_d[slot] = make_float4(dx, dy, dz, 0.f); }
        }
        float t, sdx, sdy, sdz; int gid;
        closest_hit<STAGED, true, false>(v, ox, oy, oz, dx, dy, dz, H, t, gid, sdx, sdy, sdz, n_tri, n_box);
        __stcs(p.hit + sq.base + i, make_float2(t, __int_as_float(gid)));
    }
    flush_work_counters(p, n_tri, n_box);
}

// ---- closest hit through the 4-wide BVH, warp-cooperative (k_isect_bvh, k_closest_hit_bvh).
// Traversal lengths differ wildly between the rays of a warp (a ray that leaves the mesh is done after a few nodes, one that
// grazes it visits dozens): a lane whose ray is finished takes the next ray of the pool (one atomicAdd on the pool's cursor per
// refill, for all idle lanes of the warp at once) while its neighbours keep traversing; the warp looks for idle lanes every
// BVH_BATCH node visits.
// Leaf triangles are not solved where they are found: a visit finds a leaf in a few of its 32 lanes, and the exact solve (longer
// than the four box tests) would run at those few lanes. Each leaf hit goes into a per-warp list in shared memory as
// (owner lane, triangle record); whenever 32 entries are there the warp solves them at full width: lane j takes entry j, reads the
// owner's ray from shared memory, and merges an accepted hit into the owner's result with a 64-bit atomicMin in shared memory on
// bits(|t|) << 32 | primitive << 1 | (t is -0)  -- the lexicographic minimum of (t, primitive id), which is the reference's scan
// order with strict <, so the order of the solves is free. A ray's best_t is refreshed after every solve round; between rounds
// the descent prunes with a stale (larger) bound, which only costs box tests. A lane whose traversal is finished keeps its ray
// until its queued entries are solved ("draining"); the list is emptied whenever the warp is about to refill idle lanes.
// The traversal stack lives in shared memory ((node, entry-distance key) per entry, B4_STACK entries per lane; deeper entries
// -- never seen on the bundled scenes -- spill to a local array): entries whose entry distance has meanwhile fallen behind best_t are
// dropped at pop time without a visit.
#ifndef RLPT_BVH_BATCH
#define RLPT_BVH_BATCH 8
#endif
#ifndef RLPT_BVH_REFILL
#define RLPT_BVH_REFILL 8
#endif
#ifndef RLPT_BVH_UNIFIED_POP
#define RLPT_BVH_UNIFIED_POP 0
#endif
#ifndef RLPT_BVH_FAST_PUSH
#define RLPT_BVH_FAST_PUSH 0      // branch-free pushes (unconditional stores + conditional increments): measured -0.5 % (medieval_inside), with the unified pop -2.5 %
#endif
constexpr int BVH_BATCH = RLPT_BVH_BATCH, BVH_REFILL = RLPT_BVH_REFILL;
constexpr int B4_STACK = 8;                                              // shared-memory stack entries per lane (kept small: the L1 that serves the nodes shares the SM's 256 KB with it)
constexpr int B4_SPILL = 88;                                             // local-memory overflow: 3 x depth 30 fits in 96 entries
constexpr int WQ_CAP = 32 + 32 * 4;                                      // < 32 left over + at most 4 leaves per lane and visit
// per-warp scratch in dynamic shared memory, behind the staged scene
constexpr int B4_WARP_BYTES = (B4_STACK + 1) * 32 * 8 + 32 * 8 + WQ_CAP * 4 + 6 * 32 * 4 + 16;   // stack: B4_STACK entries + one scratch slot per lane (branch-free pushes)
constexpr size_t B4_CTA_BYTES = (size_t)(BLOCK / 32) * B4_WARP_BYTES;
size_t bvh4_scratch_bytes() { return B4_CTA_BYTES; }
struct B4Scratch { uint2* stack; unsigned long long* best; unsigned* wq; float* ray; int* count; };
__device__ __forceinline__ B4Scratch b4_scratch(unsigned char* base) {
    unsigned char* p = base + (threadIdx.x >> 5) * B4_WARP_BYTES;
    B4Scratch w; w.stack = reinterpret_cast<uint2*>(p); p += (B4_STACK + 1) * 32 * 8;
    w.best = reinterpret_cast<unsigned long long*>(p); p += 32 * 8;
    w.wq = reinterpret_cast<unsigned*>(p); p += WQ_CAP * 4;
    w.ray = reinterpret_cast<float*>(p); p += 6 * 32 * 4;
    w.count = reinterpret_cast<int*>(p);
    return w;
}
// list entry = owner lane << 27 | leaf word (first record << 3 | records): the solving lane walks the leaf's 1..BVH4_LEAF_MAX records
template <bool STAGED>
__device__ __forceinline__ void wq_solve(const SceneView<STAGED>& v, const B4Scratch& w, int lo, int n, unsigned lane, float& best_t, int& first_pos, unsigned& n_tri) {
    if ((int)lane < n) {
        const unsigned e = w.wq[lo + lane];
        const int src = (int)(e >> 27), first = (int)((e & 0x7ffffffu) >> 3), cnt = (int)(e & 7u);
        const float ox = w.ray[src], oy = w.ray[32 + src], oz = w.ray[64 + src], a0 = w.ray[96 + src], a1 = w.ray[128 + src], a2 = w.ray[160 + src];
        float bt = __uint_as_float((unsigned)(w.best[src] >> 32));
        for (int j = 0; j < cnt; ++j) {
            const int pos = first + j;
            const float4 q0 = v.tri(3 * pos), q1 = v.tri(3 * pos + 1), q2 = v.tri(3 * pos + 2);
            const TriRec tr{ q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y }; const int gid = __float_as_int(q2.z);
            float t;
            n_tri++;
            if (tri_solve(tr, ox, oy, oz, a0, a1, a2, bt, t) && t < T_MISS) { atomicMin(w.best + src, hit_key(t, gid)); bt = fminf(bt, t); }
        }
    }
    __syncwarp();
    if (first_pos >= lo) first_pos = 0x7fffffff;                        // everything of mine at or above lo has been solved
    best_t = __uint_as_float((unsigned)(w.best[lane] >> 32));
}
// Src: int n (rays in the pool), fetch(i, ox, oy, oz, dx, dy, dz) -> the normalised ray i, emit(i, t, gid) <- its closest hit
template <bool STAGED, class Src>
__device__ __forceinline__ void bvh4_trace_warp(const SceneView<STAGED>& v, const Src& src, int* cursor, float H, unsigned char* scratch, unsigned& n_tri, unsigned& n_box) {
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31;
    const B4Scratch w = b4_scratch(scratch);
    uint2* stack = w.stack + lane;                                      // entry k at stack[32 k]: conflict-free
    uint2 spill[B4_SPILL];
    bool have = false, drain = false, exhausted = false;
    int i = 0, cur = 0, top = 0, count = 0, first_pos = 0x7fffffff; float best_t = T_MISS;
    Ray4 r{};
    if (lane == 0) *w.count = 0;
    __syncwarp();
    auto finish = [&]() {                                               // the ray's result: (t, primitive) from the merged key
        const unsigned long long k = w.best[lane];
        const unsigned lo = (unsigned)k;
        const float t = __uint_as_float((unsigned)(k >> 32) | (lo == 0xffffffffu ? 0u : (lo & 1u) << 31));
        src.emit(i, t, lo == 0xffffffffu ? -1 : (int)(lo >> 1));
    };
    auto push = [&](unsigned node, unsigned key) { if (top < B4_STACK) stack[32 * top] = make_uint2(node, key); else if (top < B4_STACK + B4_SPILL) spill[top - B4_STACK] = make_uint2(node, key); ++top; };
    auto solve_down_to = [&](int keep) {                                // solve rounds of up to 32 entries from the top of the list until at most `keep` are left
        while (count > keep) {
            const int n = count < 32 ? count : 32;
            count -= n;
            wq_solve<STAGED>(v, w, count, n, lane, best_t, first_pos, n_tri);
        }
        if (lane == 0) *w.count = count;
        __syncwarp();
    };
    while (true) {
        const unsigned idle = __ballot_sync(full, !have);               // idle or draining
        if (idle == full || __popc(idle) >= BVH_REFILL) {
            if (count > 0) solve_down_to(0);                            // empty the list: draining lanes become idle
            if (drain) { finish(); drain = false; }
            if (!exhausted) {
                const int leader = __ffs(idle) - 1; int base = 0;
                if ((int)lane == leader) base = atomicAdd(cursor, __popc(idle));
                base = __shfl_sync(full, base, leader);
                exhausted = base + __popc(idle) >= src.n;
                if (!have) {
                    i = base + __popc(idle & lanemask_lt());
                    if (i < src.n) {
                        float ox, oy, oz, dx, dy, dz;
                        src.fetch(i, ox, oy, oz, dx, dy, dz);
                        const float sdx = RLPT_MUL(dx, H), sdy = RLPT_MUL(dy, H), sdz = RLPT_MUL(dz, H);
                        w.ray[lane] = ox; w.ray[32 + lane] = oy; w.ray[64 + lane] = oz;
                        w.ray[96 + lane] = RLPT_SUB(0.f, sdx); w.ray[128 + lane] = RLPT_SUB(0.f, sdy); w.ray[160 + lane] = RLPT_SUB(0.f, sdz);
                        r = ray4_setup(ox, oy, oz, sdx, sdy, sdz);
                        best_t = T_MISS; w.best[lane] = KEY_MISS; cur = 0; top = 0; have = true;
                    }
                }
                __syncwarp();
            }
        }
        if (!__any_sync(full, have)) break;                             // nothing traversing: the list is empty here too
#pragma unroll 1
        for (int step = 0; step < BVH_BATCH; ++step) {
            bool leafs = false;
            if (have) {
                float tn[4]; bool hit[4]; int lk[4];
                bvh4_boxes<STAGED>(v, cur, r, best_t, tn, hit, lk);
                n_box += bvh4_children(lk);
                unsigned key[4]; bool lh[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    key[k] = (hit[k] && lk[k] >= 0) ? ((__float_as_uint(tn[k]) & ~3u) | (unsigned)k) : B4_NONE;
                    lh[k] = hit[k] && lk[k] < -1;                       // a leaf with records (BVH4_EMPTY = -1 has none)
                }
                leafs = lh[0] || lh[1] || lh[2] || lh[3];
                if (leafs) {                                            // this lane's leaf hits -> the warp's list (slots claimed with one shared-memory atomic)
                    int pos = atomicAdd(w.count, (int)lh[0] + (int)lh[1] + (int)lh[2] + (int)lh[3]);
                    first_pos = min(first_pos, pos);
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (lh[k]) w.wq[pos++] = (lane << 27) | (unsigned)(~lk[k]);
                }
                bvh4_sort(key);
#if RLPT_BVH_UNIFIED_POP
                // every inner hit goes on the stack (farthest first) and the next node is popped from it: one code path for "descend" and
                // "backtrack" instead of two that the warp's lanes would walk one after the other. Pushes are branch-free while the shared-memory
                // stack has room: four unconditional stores (slot B4_STACK is a scratch slot) and conditional increments.
                {
                    const int p0 = key[0] != B4_NONE, p1 = key[1] != B4_NONE, p2 = key[2] != B4_NONE, p3 = key[3] != B4_NONE;
                    if (top + p0 + p1 + p2 + p3 <= B4_STACK) {
                        stack[32 * top] = make_uint2((unsigned)lk[0] + (key[3] & 3u), key[3]); top += p3;
                        stack[32 * top] = make_uint2((unsigned)lk[0] + (key[2] & 3u), key[2]); top += p2;
                        stack[32 * top] = make_uint2((unsigned)lk[0] + (key[1] & 3u), key[1]); top += p1;
                        stack[32 * top] = make_uint2((unsigned)lk[0] + (key[0] & 3u), key[0]); top += p0;
                    } else {
                        if (p3) push((unsigned)lk[0] + (key[3] & 3u), key[3]);
                        if (p2) push((unsigned)lk[0] + (key[2] & 3u), key[2]);
                        if (p1) push((unsigned)lk[0] + (key[1] & 3u), key[1]);
                        if (p0) push((unsigned)lk[0] + (key[0] & 3u), key[0]);
                    }
                    bool found = false;
                    while (top > 0) {
                        --top;
                        const uint2 e = top < B4_STACK ? stack[32 * top] : spill[min(top - B4_STACK, B4_SPILL - 1)];
                        if (__uint_as_float(e.y & ~3u) <= best_t) { cur = (int)e.x; found = true; break; }
                    }
                    if (!found) { have = false; drain = true; }
                }
#else
                if (key[0] != B4_NONE) {
                    const int p1 = key[1] != B4_NONE, p2 = key[2] != B4_NONE, p3 = key[3] != B4_NONE;
                    if (RLPT_BVH_FAST_PUSH && top + p1 + p2 + p3 <= B4_STACK) {     // branch-free: unconditional stores (slot B4_STACK is scratch), conditional increments
                        stack[32 * top] = make_uint2((unsigned)lk[0] + (key[3] & 3u), key[3]); top += p3;
                        stack[32 * top] = make_uint2((unsigned)lk[0] + (key[2] & 3u), key[2]); top += p2;
                        stack[32 * top] = make_uint2((unsigned)lk[0] + (key[1] & 3u), key[1]); top += p1;
                    } else {
                        if (p3) push((unsigned)lk[0] + (key[3] & 3u), key[3]);
                        if (p2) push((unsigned)lk[0] + (key[2] & 3u), key[2]);
                        if (p1) push((unsigned)lk[0] + (key[1] & 3u), key[1]);
                    }
                    cur = lk[0] + (int)(key[0] & 3u);
                } else {
                    bool found = false;
                    while (top > 0) {
                        --top;
                        const uint2 e = top < B4_STACK ? stack[32 * top] : spill[min(top - B4_STACK, B4_SPILL - 1)];
                        if (__uint_as_float(e.y & ~3u) <= best_t) { cur = (int)e.x; found = true; break; }
                    }
                    if (!found) { have = false; drain = true; }
                }
#endif
            }
            if (__any_sync(full, leafs)) {
                __syncwarp();                                           // the list writes above are visible to the solving lanes
                count = *w.count;
                if (count >= 32) solve_down_to(31);
            }
            if (drain && first_pos == 0x7fffffff) { finish(); drain = false; }
        }
    }
}

template <bool PRIMARY>
struct QueueRays {
    const FrameParams& p; const FrameDyn& dyn; PathQueue qi; int base, n, bounce;
    __device__ __forceinline__ void fetch(int i, float& ox, float& oy, float& oz, float& dx, float& dy, float& dz) const {
        if (PRIMARY) { PathState s; primary_state(p, dyn, base + i, s); ox = s.ox; oy = s.oy; oz = s.oz; dx = s.dx; dy = s.dy; dz = s.dz; }
        else { const float4 a = __ldcs(qi.o + base + i), b = __ldcs(qi.d + base + i); ox = a.x; oy = a.y; oz = a.z; dx = b.x; dy = b.y; dz = b.z; }
        if (dyn.capture_max > 0 && bounce == dyn.capture_bounce) {
            const int slot = atomicAdd(p.capture_n, 1);
            if (slot < dyn.capture_max) { p.capture_o[slot] = make_float4(ox, oy, oz, 0.f); p.capture_d[slot] = make_float4(dx, dy, dz, 0.f); }
        }
    }
    __device__ __forceinline__ void emit(int i, float t, int gid) const { __stcs(p.hit + base + i, make_float2(t, __int_as_float(gid))); }
};
#ifndef RLPT_BVH_MINBLOCKS
#define RLPT_BVH_MINBLOCKS 4
#endif
template <bool STAGED, bool PRIMARY>
__global__ void __launch_bounds__(BLOCK, RLPT_BVH_MINBLOCKS) k_isect_bvh(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, int bounce) {
    const SubQueue sq = sub_queue<PRIMARY>(p, bounce);
    if ((int)(blockIdx.x / NSUB) * BLOCK >= sq.n) return;
    SceneView<STAGED> v = stage_scene<STAGED, false>(p.scene);
    unsigned char* scratch = reinterpret_cast<unsigned char*>(s_scene) + (STAGED ? scene_smem_bytes_dev(p.scene) : 0);
    int* cursor = p.cursor + (bounce * NSUB + (int)(blockIdx.x % NSUB)) * COUNT_STRIDE;
    unsigned n_tri = 0, n_box = 0;
    const QueueRays<PRIMARY> src{ p, dyn, p.q[bounce & 1], sq.base, sq.n, bounce };
    bvh4_trace_warp<STAGED>(v, src, cursor, (float)p.height, scratch, n_tri, n_box);
    flush_work_counters(p, n_tri, n_box);
}

// the parity entry point rlpt_closest_hit through the same warp-cooperative traversal (scenes walked through the BVH)
struct ArrayRays {
    const float* org; const float* dir; int n, n_surf; int* type; int* index; float* t_out;
    __device__ __forceinline__ void fetch(int i, float& ox, float& oy, float& oz, float& dx, float& dy, float& dz) const {
        const f3 d = normalize_ref(f3{ dir[3 * i], dir[3 * i + 1], dir[3 * i + 2] });
        ox = org[3 * i]; oy = org[3 * i + 1]; oz = org[3 * i + 2]; dx = d.x; dy = d.y; dz = d.z;
    }
    __device__ __forceinline__ void emit(int i, float t, int gid) const {
        type[i] = gid < 0 ? 0 : (gid < n_surf ? 2 : 1);
        index[i] = gid < 0 ? -1 : (gid < n_surf ? gid : gid - n_surf);
        t_out[i] = t;
    }
};
template <bool STAGED>
__global__ void __launch_bounds__(BLOCK, 3) k_closest_hit_bvh(SceneDev sc, const float* __restrict__ org, const float* __restrict__ dir, int n, float H,
                                                              int* __restrict__ type, int* __restrict__ index, float* __restrict__ t_out,
                                                              unsigned long long* __restrict__ counters, int* __restrict__ cursor) {
    SceneView<STAGED> v = stage_scene<STAGED, false>(sc);
    unsigned char* scratch = reinterpret_cast<unsigned char*>(s_scene) + (STAGED ? scene_smem_bytes_dev(sc) : 0);
    unsigned nt = 0, nb = 0;
    const ArrayRays src{ org, dir, n, sc.n_surf, type, index, t_out };
    bvh4_trace_warp<STAGED>(v, src, cursor, H, scratch, nt, nb);
    if (counters) {
        nt = __reduce_add_sync(0xffffffffu, nt); nb = __reduce_add_sync(0xffffffffu, nb);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&counters[0], (unsigned long long)nt); atomicAdd(&counters[1], (unsigned long long)nb); }
    }
}

void launch_closest_hit(const SceneDev& sc, const float* org, const float* dir, int n, float H, int* type, int* index, float* t,
                        unsigned long long* counters, int* cursor, size_t smem, cudaStream_t s) {
    int grid = (n + BLOCK - 1) / BLOCK; if (grid > 148 * 8) grid = 148 * 8; if (grid < 1) grid = 1;
    const bool staged = sc.staged != 0;
    if (!sc.brute) {
        // warp-cooperative traversal (the code k_isect_bvh runs): rays are handed out through `cursor` (zeroed here)
        cudaMemsetAsync(cursor, 0, sizeof(int), s);
        if (grid > 148 * 3) grid = 148 * 3;
        if (staged) k_closest_hit_bvh<true><<<grid, BLOCK, smem + B4_CTA_BYTES, s>>>(sc, org, dir, n, H, type, index, t, counters, cursor);
        else k_closest_hit_bvh<false><<<grid, BLOCK, B4_CTA_BYTES, s>>>(sc, org, dir, n, H, type, index, t, counters, cursor);
        return;
    }
    if (staged) {
        if (counters) k_closest_hit<true, true><<<grid, BLOCK, smem, s>>>(sc, org, dir, n, H, type, index, t, counters);
        else k_closest_hit<true, false><<<grid, BLOCK, smem, s>>>(sc, org, dir, n, H, type, index, t, counters);
    } else {
        if (counters) k_closest_hit<false, true><<<grid, BLOCK, smem, s>>>(sc, org, dir, n, H, type, index, t, counters);
        else k_closest_hit<false, false><<<grid, BLOCK, smem, s>>>(sc, org, dir, n, H, type, index, t, counters);
    }
}

// Split, second half: shading records come through the read-only path (2.4 KB for Cornell: L1-resident), no staging
template <bool SARSA, bool PRIMARY>
__global__ void __launch_bounds__(BLOCK, RLPT_SHADE_MINBLOCKS) k_shade(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, int bounce) {
    const SubQueue sq = sub_queue<PRIMARY>(p, bounce);
    const int n_round = (sq.n + 31) & ~31;
    const PathQueue qi = p.q[bounce & 1], qo = p.q[(bounce + 1) & 1];
    unsigned st_len = 0, st_zero = 0, st_term = 0, n_kd = 0;
    const float4* __restrict__ shade_g = p.scene.shade;
    auto shade = [&](int k) { return __ldg(shade_g + k); };
    for (int i = sq.first; i < n_round; i += sq.stride) {
        bool live = i < sq.n;
        PathState s{}; float t = T_MISS; int gid = -1;
        if (live) {
            if (PRIMARY) {
                primary_state(p, dyn, sq.base + i, s);
                if ((sq.base + i) % p.spp == 0) atomicAdd(&p.accum[s.pixel].w, (float)p.spp);
            } else load_state(qi, sq.base + i, s);
            float2 h = __ldcs(p.hit + sq.base + i); t = h.x; gid = __float_as_int(h.y);
        }
        live = shade_step<SARSA, !PRIMARY>(p, dyn, shade, p.scene.n_surf, bounce, live, s, t, gid, st_len, st_zero, st_term, n_kd);
        compact_store(sq, qo, bounce, live, s);
    }
    flush_path_stats(p, st_len, st_zero, st_term, n_kd);
}

// Debug view of the radiance map: draw_voronoi_trace (G/path_tracing/voronoi_trace.cu:4-45): one jittered camera ray per
// pixel; a surface hit is painted with the colour of its nearest radiance volume, anything else white. The reference gives
// every volume a host rand() colour (radiance_volume.cu:311-318); here the colour is Philox(seed, volume) -- arbitrary in
// both. The frame buffer is SET (weight 1), not accumulated.
template <bool STAGED>
__global__ void __launch_bounds__(BLOCK) k_voronoi(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn) {
    SceneView<STAGED> v = stage_scene<STAGED, true>(p.scene);
    const int n = p.width * p.height;
    unsigned n_tri = 0, n_box = 0, n_kd = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        PathState s; primary_state(p, dyn, i, s);                       // p.spp == 1 here: path index = pixel
        float t, sdx, sdy, sdz; int gid;
        closest_hit<STAGED, true>(v, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, (float)p.height, t, gid, sdx, sdy, sdz, n_tri, n_box);
        float4 c = make_float4(1.f, 1.f, 1.f, 1.f);
        if (gid >= 0 && gid < v.n_surf) {
            const float hx = RLPT_FMA(sdx, t, s.ox), hy = RLPT_FMA(sdy, t, s.oy), hz = RLPT_FMA(sdz, t, s.oz);
            const int nv = find_volume(p.rm, hx, hy, hz, __float_as_int(v.shade(4 * gid + 1).w), n_kd);
            float u0, u1, u2, u3; draw4(p.seed, (uint32_t)nv, 0x766f726fu, 0u, 7u, u0, u1, u2, u3);
            c = make_float4(u0, u1, u2, 1.f);
        }
        p.accum[i] = c;
    }
}
void launch_voronoi(const FrameParams& p, const FrameDyn& dyn, int grid, size_t smem, cudaStream_t s) {
    const SceneDev& sc = p.scene;
    if (sc.staged) k_voronoi<true><<<grid, BLOCK, smem, s>>>(p, dyn);
    else k_voronoi<false><<<grid, BLOCK, smem, s>>>(p, dyn);
}

static bool scene_staged(const SceneDev& sc) { return sc.staged != 0; }
template <bool SARSA, bool PRIMARY, bool TAIL>
static void launch_bounce_t(const FrameParams& p, const FrameDyn& dyn, int bounce, int grid, size_t smem, cudaStream_t s) {
    if (scene_staged(p.scene)) k_bounce<true, SARSA, PRIMARY, TAIL><<<grid, BLOCK, smem, s>>>(p, dyn, bounce);
    else k_bounce<false, SARSA, PRIMARY, TAIL><<<grid, BLOCK, smem, s>>>(p, dyn, bounce);
}
void launch_primary(const FrameParams& p, const FrameDyn& dyn, int method, int grid, size_t smem, cudaStream_t s) {
    if (method == 1) launch_bounce_t<true, true, false>(p, dyn, 0, grid, smem, s); else launch_bounce_t<false, true, false>(p, dyn, 0, grid, smem, s);
}
void launch_bounce(const FrameParams& p, const FrameDyn& dyn, int method, int bounce, int grid, size_t smem, cudaStream_t s) {
    if (method == 1) launch_bounce_t<true, false, false>(p, dyn, bounce, grid, smem, s); else launch_bounce_t<false, false, false>(p, dyn, bounce, grid, smem, s);
}
void launch_tail(const FrameParams& p, const FrameDyn& dyn, int method, int bounce, int grid, size_t smem, cudaStream_t s) {
    if (method == 1) launch_bounce_t<true, false, true>(p, dyn, bounce, grid, smem, s); else launch_bounce_t<false, false, true>(p, dyn, bounce, grid, smem, s);
}
void launch_isect(const FrameParams& p, const FrameDyn& dyn, int bounce, int grid, size_t smem, cudaStream_t s) {
    const bool staged = scene_staged(p.scene);
    if (!p.scene.brute) {                                                  // (p.cursor is set by every caller: the frame loops allocate it with the queues)
        const size_t sm = smem + B4_CTA_BYTES;
        if (bounce == 0) { if (staged) k_isect_bvh<true, true><<<grid, BLOCK, sm, s>>>(p, dyn, 0); else k_isect_bvh<false, true><<<grid, BLOCK, sm, s>>>(p, dyn, 0); }
        else { if (staged) k_isect_bvh<true, false><<<grid, BLOCK, sm, s>>>(p, dyn, bounce); else k_isect_bvh<false, false><<<grid, BLOCK, sm, s>>>(p, dyn, bounce); }
        return;
    }
    if (bounce == 0) { if (staged) k_isect<true, true><<<grid, BLOCK, smem, s>>>(p, dyn, 0); else k_isect<false, true><<<grid, BLOCK, smem, s>>>(p, dyn, 0); }
    else { if (staged) k_isect<true, false><<<grid, BLOCK, smem, s>>>(p, dyn, bounce); else k_isect<false, false><<<grid, BLOCK, smem, s>>>(p, dyn, bounce); }
}
void launch_shade(const FrameParams& p, const FrameDyn& dyn, int method, int bounce, int grid, cudaStream_t s) {
    if (method == 1) { if (bounce == 0) k_shade<true, true><<<grid, BLOCK, 0, s>>>(p, dyn, 0); else k_shade<true, false><<<grid, BLOCK, 0, s>>>(p, dyn, bounce); }
    else { if (bounce == 0) k_shade<false, true><<<grid, BLOCK, 0, s>>>(p, dyn, 0); else k_shade<false, false><<<grid, BLOCK, 0, s>>>(p, dyn, bounce); }
}

// ------------------------------------------------------------------------------------------------ Neural-Q wavefront
// PretrainedPathtracer (G/deep_learning/pre_trained_pathtracer.cu:186-491) as a compacted wavefront: per bounce
//   k_nq_trace  closest hit of every live path (initialise_ray :379-418 fused into bounce 0), throughput *= light power
//               / diffuse_c/pi (trace_ray :420-491), terminated paths go to the frame buffer, survivors are compacted with
//               their hit point (the network's query point) and surface id
//   k_dqn_forward over the survivors' hit points (rlpt_dqn.cu)   [replaces :228-271]
//   k_nq_sample importance-samples the next direction from Q(s, .) cos(theta) (importance_sample_direction,
//               nn_rendering_helpers.cu:391-489) and applies cos(theta)/pdf to the throughput
enum { PURPOSE_NQ = 2 };
template <bool STAGED, bool PRIMARY>
__global__ void __launch_bounds__(BLOCK) k_nq_trace(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, int bounce) {
    SceneView<STAGED> v = stage_scene<STAGED>(p.scene);
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31;
    const PathQueue qi = p.q[bounce & 1], qo = p.q[(bounce + 1) & 1];
    const int n_in = PRIMARY ? p.width * p.height * p.spp : p.counts[bounce];
    unsigned st_len = 0, st_zero = 0, st_term = 0, n_tri = 0, n_box = 0;
    const float H = (float)p.height;
    const int n_round = (n_in + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n_in;
        float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1, tr = 1, tg = 1, tb = 1; uint32_t pixel = 0, sample = 0;
        if (valid) {
            if (PRIMARY) {
                pixel = (uint32_t)(i / p.spp); sample = dyn.sample_base + (uint32_t)(i % p.spp);
                float u0, u1, u2, u3; draw4(p.seed, pixel, sample, 0u, PURPOSE_CAMERA, u0, u1, u2, u3);
                f3 d = camera_dir((int)(pixel / (uint32_t)p.height), (int)(pixel % (uint32_t)p.height), u0, u1, p.width, p.height, dyn.rotated != 0, dyn.cy, dyn.sy, dyn.cx, dyn.sx);
                ox = dyn.cam_x; oy = dyn.cam_y; oz = dyn.cam_z; dx = d.x; dy = d.y; dz = d.z;
                if (i % p.spp == 0) atomicAdd(&p.accum[pixel].w, (float)p.spp);
            } else {
                float4 a = qi.o[i], b = qi.d[i], c = qi.thr[i]; uint32_t m = qi.meta[i];
                pixel = __float_as_uint(a.w); sample = m; tr = c.x; tg = c.y; tb = c.z;
                ox = RLPT_FMA(RAY_EPS, b.x, a.x); oy = RLPT_FMA(RAY_EPS, b.y, a.y); oz = RLPT_FMA(RAY_EPS, b.z, a.z);      // position + dir * 0.00001f (:439)
                f3 nn = normalize_ref(f3{ b.x, b.y, b.z }); dx = nn.x; dy = nn.y; dz = nn.z;
            }
        }
        float t = T_MISS, sdx = 0, sdy = 0, sdz = 0; int gid = -1;
        if (valid) closest_hit<STAGED, true>(v, ox, oy, oz, dx, dy, dz, H, t, gid, sdx, sdy, sdz, n_tri, n_box);
        const bool surface = valid && gid >= 0 && gid < v.n_surf;
        bool alive = false; float hx = 0, hy = 0, hz = 0;
        if (valid && !surface) {
            float lr = p.env, lg = p.env, lb = p.env;
            if (gid >= v.n_surf) { float4 e = v.shade(4 * gid + 3); lr = e.x; lg = e.y; lb = e.z; }
            lr *= tr; lg *= tg; lb *= tb;
            if (lr != 0.f || lg != 0.f || lb != 0.f) atomicAdd(p.accum + pixel, make_float4(lr, lg, lb, 0.f));
            st_len += (unsigned)bounce + 1u; st_term++;
            if (zero_contribution(lr, lg, lb)) st_zero++;
        } else if (surface) {
            if (bounce + 1 >= p.max_bounces) { st_len += (unsigned)p.max_bounces; st_term++; st_zero++; }     // out of bounces: contributes nothing (DESIGN.md deviation 8)
            else {
                hx = RLPT_FMA(sdx, t, ox); hy = RLPT_FMA(sdy, t, oy); hz = RLPT_FMA(sdz, t, oz);
                float4 sC = v.shade(4 * gid + 3); tr *= sC.x; tg *= sC.y; tb *= sC.z;                      // BRDF = diffuse_c / pi (:476-480)
                alive = true;
            }
        }
        unsigned bal = __ballot_sync(full, alive);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(p.counts + bounce + 1, __popc(bal));
            base = __shfl_sync(full, base, 0);
            if (alive) {
                int slot = base + __popc(bal & lanemask_lt());
                qo.o[slot] = make_float4(hx, hy, hz, __uint_as_float(pixel));
                qo.d[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(gid));
                qo.thr[slot] = make_float4(tr, tg, tb, 0.f);
                qo.meta[slot] = sample;
            }
        }
    }
    st_len = __reduce_add_sync(full, st_len); st_zero = __reduce_add_sync(full, st_zero); st_term = __reduce_add_sync(full, st_term);
    if (lane == 0 && st_term) { atomicAdd(p.stats + 0, (unsigned long long)st_len); atomicAdd(p.stats + 1, (unsigned long long)st_zero); atomicAdd(p.stats + 2, (unsigned long long)st_term); }
    n_tri = __reduce_add_sync(full, n_tri); n_box = __reduce_add_sync(full, n_box);
    if (lane == 0 && (n_tri | n_box)) { atomicAdd(p.stats + 3, (unsigned long long)n_tri); atomicAdd(p.stats + 4, (unsigned long long)n_box); }
}
void launch_nq_trace(const FrameParams& p, const FrameDyn& dyn, int bounce, int grid, size_t smem, cudaStream_t s) {
    const SceneDev& sc = p.scene;
    bool staged = sc.staged != 0;
    if (bounce == 0) { if (staged) k_nq_trace<true, true><<<grid, BLOCK, smem, s>>>(p, dyn, 0); else k_nq_trace<false, true><<<grid, BLOCK, smem, s>>>(p, dyn, 0); }
    else { if (staged) k_nq_trace<true, false><<<grid, BLOCK, smem, s>>>(p, dyn, bounce); else k_nq_trace<false, false><<<grid, BLOCK, smem, s>>>(p, dyn, bounce); }
}

// One thread per live path of queue `bounce`. q: [144][q_stride] from k_dqn_forward. Weighting uses the cell-centre cosine
// table (the reference draws a fresh random point in every cell just to weight it, :418-437; DESIGN.md deviation 7).
// epsilon > 0: epsilon-greedy exploration (sample_batch_ray_directions_epsilon_greedy, :330-389); action_out (may be null)
// receives the chosen cell per queue slot.
__global__ void __launch_bounds__(BLOCK) k_nq_sample(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, int bounce,
                                                     const float* __restrict__ q, int q_stride, float epsilon, uint32_t* __restrict__ action_out) {
    const PathQueue qu = p.q[bounce & 1];
    const int n = p.counts[bounce];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 a = qu.o[i], b = qu.d[i], c = qu.thr[i]; const uint32_t m = qu.meta[i];
        const int gid = __float_as_int(b.w); const uint32_t pixel = __float_as_uint(a.w), sample = m;
        float4 sN = __ldg(p.scene.shade + 4 * gid), sT = __ldg(p.scene.shade + 4 * gid + 1), sB = __ldg(p.scene.shade + 4 * gid + 2);
        f3 N = { sN.x, sN.y, sN.z }, T = { sT.x, sT.y, sT.z }, B = { sB.x, sB.y, sB.z };
        float u0, u1, u2, u3; draw4(p.seed, pixel, sample, (uint32_t)bounce, PURPOSE_NQ, u0, u1, u2, u3);
        int cell; float pdf;
        if (dyn.max_dir) {
            // sample_max_direction (nn_rendering_helpers.cu:492-553): the first cell with the largest Q (cell 0 when none is positive), a random point
            // inside it, and the reference's constant pdf 2 RHO
            cell = 0; float mq = 0.f;
            for (int k = 0; k < CELLS; ++k) { const float qk = __ldg(q + (size_t)k * q_stride + i); if (qk > mq) { mq = qk; cell = k; } }
            pdf = RHO * 2.f;
        } else if (u3 <= epsilon) {                                            // explore: a uniformly chosen cell, pdf = RHO (:366-387, :30-36)
            cell = min((int)(u0 * (float)CELLS), CELLS - 1); pdf = RHO;
        } else {
            float total = 0.f;
            for (int k = 0; k < CELLS; ++k) total += __ldg(q + (size_t)k * q_stride + i) * c_cell_cos[k];
            const bool dead = !(total > 0.f);                                 // ReLU on the output layer can zero a whole row: fall back to cosine weighting
            if (dead) { total = 0.f; for (int k = 0; k < CELLS; ++k) total += c_cell_cos[k]; }
            const float r = u0 * total; float run = 0.f, w = 0.f; cell = -1; int last = 0; float last_w = 0.f;
            for (int k = 0; k < CELLS; ++k) {
                w = (dead ? 1.f : __ldg(q + (size_t)k * q_stride + i)) * c_cell_cos[k]; run += w;
                if (w > 0.f) { last = k; last_w = w; }
                if (run > r && w > 0.f) { cell = k; break; }
            }
            if (cell < 0) { cell = last; w = last_w; }                          // r == total after rounding: the last cell with weight
            pdf = RHO * ((w / total) / GRID_RHO);
        }
        f3 nd = grid_to_direction((float)(cell / GRID) + u1, (float)(cell % GRID) + u2, T, N, B);
        const float cos_theta = N.x * nd.x + N.y * nd.y + N.z * nd.z;
        const float scale = cos_theta / pdf;
        qu.d[i] = make_float4(nd.x, nd.y, nd.z, b.w);
        qu.thr[i] = make_float4(c.x * scale, c.y * scale, c.z * scale, 0.f);
        if (action_out) action_out[i] = (uint32_t)cell;
    }
}
void launch_nq_sample(const FrameParams& p, const FrameDyn& dyn, int bounce, const float* q, int q_stride, float epsilon, uint32_t* action_out, int grid, cudaStream_t s) {
    k_nq_sample<<<grid, BLOCK, 0, s>>>(p, dyn, bounce, q, q_stride, epsilon, action_out);
}

// ------------------------------------------------------------------------------------------------ Neural-Q training tracer
// NeuralQPathtracer::render_frame (G/deep_learning/neural_q_pathtracer.cu:226-600) without the host round trips: the same
// per-bounce sequence (sample with epsilon-greedy -> trace all rays -> TD targets -> optimiser steps -> re-seed terminated
// rays), state in SoA float4 arrays, the network evaluated by k_dqn_forward on the rays' positions directly.
enum { PURPOSE_NQ_RESPAWN = 3 };
__global__ void k_nqt_init(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, NqTrainState st) {      // initialise_ray (:603-643)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    float u0, u1, u2, u3; draw4(p.seed, (uint32_t)i, dyn.sample_base, 0u, PURPOSE_CAMERA, u0, u1, u2, u3);
    f3 d = camera_dir(i / p.height, i % p.height, u0, u1, p.width, p.height, dyn.rotated != 0, dyn.cy, dyn.sy, dyn.cx, dyn.sx);
    st.loc[i] = make_float4(dyn.cam_x, dyn.cam_y, dyn.cam_z, __int_as_float(-1)); st.sloc[i] = st.loc[i];
    st.dir[i] = make_float4(d.x, d.y, d.z, 0.f); st.thr[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    st.state[i] = 0u; st.reward[i] = 0.f; st.discount[i] = 1.f; st.action[i] = 0u;
    atomicAdd(&p.accum[i].w, 1.f);
    if (i == 0) st.alive[0] = st.n;
}
void launch_nqt_init(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, cudaStream_t s) { k_nqt_init<<<(st.n + 255) / 256, 256, 0, s>>>(p, dyn, st); }

// sample_batch_ray_directions_epsilon_greedy (nn_rendering_helpers.cu:330-389) for every ray; only alive rays carry throughput
__global__ void __launch_bounds__(BLOCK) k_nqt_sample(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, NqTrainState st, int bounce,
                                                      const float* __restrict__ q, int q_stride, float epsilon) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    const float4 a = st.loc[i]; const int gid = __float_as_int(a.w);
    st.sloc[i] = a;
    if (gid < 0) return;                                                       // cannot happen after bounce 0 (terminated rays are re-seeded on a surface)
    float4 sN = __ldg(p.scene.shade + 4 * gid), sT = __ldg(p.scene.shade + 4 * gid + 1), sB = __ldg(p.scene.shade + 4 * gid + 2);
    f3 N = { sN.x, sN.y, sN.z }, T = { sT.x, sT.y, sT.z }, B = { sB.x, sB.y, sB.z };
    float u0, u1, u2, u3; draw4(p.seed, (uint32_t)i, dyn.sample_base, (uint32_t)bounce, PURPOSE_NQ, u0, u1, u2, u3);
    int cell; float pdf;
    if (u3 <= epsilon) { cell = min((int)(u0 * (float)CELLS), CELLS - 1); pdf = RHO; }
    else {
        float total = 0.f;
        for (int k = 0; k < CELLS; ++k) total += __ldg(q + (size_t)k * q_stride + i) * c_cell_cos[k];
        const bool dead = !(total > 0.f);
        if (dead) { total = 0.f; for (int k = 0; k < CELLS; ++k) total += c_cell_cos[k]; }
        const float r = u0 * total; float run = 0.f, w = 0.f; cell = -1; int last = 0; float last_w = 0.f;
        for (int k = 0; k < CELLS; ++k) {
            w = (dead ? 1.f : __ldg(q + (size_t)k * q_stride + i)) * c_cell_cos[k]; run += w;
            if (w > 0.f) { last = k; last_w = w; }
            if (run > r && w > 0.f) { cell = k; break; }
        }
        if (cell < 0) { cell = last; w = last_w; }
        pdf = RHO * ((w / total) / GRID_RHO);
    }
    f3 nd = grid_to_direction((float)(cell / GRID) + u1, (float)(cell % GRID) + u2, T, N, B);
    st.dir[i] = make_float4(nd.x, nd.y, nd.z, 0.f); st.action[i] = (uint32_t)cell;
    if (st.state[i] == 0u) {
        const float scale = (N.x * nd.x + N.y * nd.y + N.z * nd.z) / pdf;
        float4 c = st.thr[i]; st.thr[i] = make_float4(c.x * scale, c.y * scale, c.z * scale, 0.f);
    }
}
void launch_nqt_sample(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, int bounce, const float* q, int q_stride, float epsilon, cudaStream_t s) {
    k_nqt_sample<<<(st.n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(p, dyn, st, bounce, q, q_stride, epsilon);
}

// trace_ray (neural_q_pathtracer.cu:646-752): every ray, alive or learning-only
template <bool STAGED>
__global__ void __launch_bounds__(BLOCK) k_nqt_trace(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, NqTrainState st, int bounce) {
    SceneView<STAGED> v = stage_scene<STAGED>(p.scene);
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31;
    unsigned st_len = 0, st_zero = 0, st_term = 0, n_tri = 0, n_box = 0, n_alive = 0;
    const float H = (float)p.height;
    const int n_round = (st.n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < st.n;
        if (valid) {
            const float4 a = st.loc[i], b = st.dir[i]; const uint32_t state = st.state[i];
            const float ox = RLPT_FMA(RAY_EPS, b.x, a.x), oy = RLPT_FMA(RAY_EPS, b.y, a.y), oz = RLPT_FMA(RAY_EPS, b.z, a.z);
            const f3 nn = normalize_ref(f3{ b.x, b.y, b.z });
            float t, sdx, sdy, sdz; int gid;
            closest_hit<STAGED, true>(v, ox, oy, oz, nn.x, nn.y, nn.z, H, t, gid, sdx, sdy, sdz, n_tri, n_box);
            float4 c = st.thr[i];
            if (gid < 0 || gid >= v.n_surf) {                                  // NOTHING / AREA_LIGHT: terminal
                float lr = p.env, lg = p.env, lb = p.env, reward = 0.f;
                if (gid >= 0) { float4 e = v.shade(4 * gid + 3); lr = e.x; lg = e.y; lb = e.z; reward = v.shade(4 * gid).w * 200.f; }     // luminance * 200 (:693)
                st.reward[i] = reward; st.discount[i] = 0.f;
                if (state == 0u) {
                    lr *= c.x; lg *= c.y; lb *= c.z;
                    st.thr[i] = make_float4(lr, lg, lb, 0.f);
                    if (lr != 0.f || lg != 0.f || lb != 0.f) atomicAdd(p.accum + i, make_float4(lr, lg, lb, 0.f));
                    st_len += (unsigned)bounce + 1u; st_term++;
                    if (zero_contribution(lr, lg, lb)) st_zero++;
                }
                st.state[i] = 1u;
            } else {                                                           // SURFACE
                const float4 sC = v.shade(4 * gid + 3);
                st.loc[i] = make_float4(RLPT_FMA(sdx, t, ox), RLPT_FMA(sdy, t, oy), RLPT_FMA(sdz, t, oz), __int_as_float(gid));
                st.reward[i] = 0.f; st.discount[i] = sC.w;                     // luminance of the surface (:724-737)
                if (state == 0u) {
                    if (bounce + 1 >= p.max_bounces) { st_len += (unsigned)p.max_bounces; st_term++; st_zero++; st.state[i] = 2u; }     // out of bounces: contributes nothing
                    else { st.thr[i] = make_float4(c.x * sC.x, c.y * sC.y, c.z * sC.z, 0.f); n_alive++; }
                }
            }
        }
    }
    n_alive = __reduce_add_sync(full, n_alive);
    if (lane == 0 && n_alive) atomicAdd(st.alive + bounce + 1, (int)n_alive);
    st_len = __reduce_add_sync(full, st_len); st_zero = __reduce_add_sync(full, st_zero); st_term = __reduce_add_sync(full, st_term);
    if (lane == 0 && st_term) { atomicAdd(p.stats + 0, (unsigned long long)st_len); atomicAdd(p.stats + 1, (unsigned long long)st_zero); atomicAdd(p.stats + 2, (unsigned long long)st_term); }
    n_tri = __reduce_add_sync(full, n_tri); n_box = __reduce_add_sync(full, n_box);
    if (lane == 0 && (n_tri | n_box)) { atomicAdd(p.stats + 3, (unsigned long long)n_tri); atomicAdd(p.stats + 4, (unsigned long long)n_box); }
}
void launch_nqt_trace(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, int bounce, int grid, size_t smem, cudaStream_t s) {
    const SceneDev& sc = p.scene;
    bool staged = sc.staged != 0;
    if (staged) k_nqt_trace<true><<<grid, BLOCK, smem, s>>>(p, dyn, st, bounce); else k_nqt_trace<false><<<grid, BLOCK, smem, s>>>(p, dyn, st, bounce);
}

// (compute_td_targets, nn_rendering_helpers.cu:91-140 -- reward + discount * max_a Q(s', a) cos(theta_a); terminal: the reward -- is evaluated inside the
// training step's backward kernel, rlpt_dqn.cu: k_dqn_backward phase 0 / k_delta3; the batch's slice of the ray arrays is read in place)

// sample_random_scene_pos_for_terminated_rays (nn_rendering_helpers.cu:241-277): a uniformly chosen surface, a uniform point
// on it. Fixed here (DESIGN.md deviation 9): the index cannot reach n_surfaces, and y / z are not swapped.
__global__ void k_nqt_respawn(const __grid_constant__ FrameParams p, const __grid_constant__ FrameDyn dyn, NqTrainState st, int bounce) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n || st.state[i] != 1u) return;
    float u0, u1, u2, u3; draw4(p.seed, (uint32_t)i, dyn.sample_base, (uint32_t)bounce, PURPOSE_NQ_RESPAWN, u0, u1, u2, u3);
    const int gid = min((int)(u0 * (float)p.scene.n_surf), p.scene.n_surf - 1);
    if (u1 + u2 > 1.f) { u1 = 1.f - u1; u2 = 1.f - u2; }
    const float4 t0 = __ldg(p.scene.tri + 3 * gid), t1 = __ldg(p.scene.tri + 3 * gid + 1), t2 = __ldg(p.scene.tri + 3 * gid + 2);
    st.loc[i] = make_float4(t0.x + u1 * t0.w + u2 * t1.z, t0.y + u1 * t1.x + u2 * t1.w, t0.z + u1 * t1.y + u2 * t2.x, __int_as_float(gid));
    st.state[i] = 2u;
}
void launch_nqt_respawn(const FrameParams& p, const FrameDyn& dyn, const NqTrainState& st, int bounce, cudaStream_t s) { k_nqt_respawn<<<(st.n + 255) / 256, 256, 0, s>>>(p, dyn, st, bounce); }
__global__ void k_add_scalar(float* dst, const float* src) { *dst += *src; }
void launch_add_scalar(float* dst, const float* src, cudaStream_t s) { k_add_scalar<<<1, 1, 0, s>>>(dst, src); }

// CTAs of k_isect / k_shade that are resident per SM (occupancy API): the split pipeline launches exactly one wave of each,
// so no CTA starts when most of the grid has already finished (+4.8 % over 8 CTAs per SM for both)
void kernels_resident_ctas(size_t isect_smem, int brute, int staged, int* isect_per_sm, int* shade_per_sm) {
    int a = 0, b = 0;
    cudaError_t e;
    if (brute) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_isect<true, false>, BLOCK, isect_smem);
    else if (staged) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_isect_bvh<true, false>, BLOCK, isect_smem + B4_CTA_BYTES);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_isect_bvh<false, false>, BLOCK, isect_smem + B4_CTA_BYTES);
    if (e != cudaSuccess) a = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_shade<true, false>, BLOCK, 0) != cudaSuccess) b = 0;
    (void)cudaGetLastError();
    *isect_per_sm = a > 0 ? a : 6; *shade_per_sm = b > 0 ? b : 4;
}

int kernels_set_smem_limit(size_t bytes) {
    cudaError_t e = cudaSuccess;
#define RLPT_SET(k) if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)
    RLPT_SET((k_bounce<true, true, true, false>)); RLPT_SET((k_bounce<true, true, false, false>)); RLPT_SET((k_bounce<true, false, true, false>)); RLPT_SET((k_bounce<true, false, false, false>));
    RLPT_SET((k_bounce<false, true, true, false>)); RLPT_SET((k_bounce<false, true, false, false>)); RLPT_SET((k_bounce<false, false, true, false>)); RLPT_SET((k_bounce<false, false, false, false>));
    RLPT_SET((k_bounce<true, true, false, true>)); RLPT_SET((k_bounce<true, false, false, true>)); RLPT_SET((k_bounce<false, true, false, true>)); RLPT_SET((k_bounce<false, false, false, true>));
    RLPT_SET((k_isect<true, true>)); RLPT_SET((k_isect<true, false>)); RLPT_SET((k_isect<false, true>)); RLPT_SET((k_isect<false, false>));
    RLPT_SET((k_isect_bvh<true, true>)); RLPT_SET((k_isect_bvh<true, false>)); RLPT_SET((k_isect_bvh<false, true>)); RLPT_SET((k_isect_bvh<false, false>));
    RLPT_SET((k_nqt_trace<true>)); RLPT_SET((k_nqt_trace<false>));
    RLPT_SET((k_nq_trace<true, true>)); RLPT_SET((k_nq_trace<true, false>)); RLPT_SET((k_nq_trace<false, true>)); RLPT_SET((k_nq_trace<false, false>));
    RLPT_SET((k_voronoi<true>)); RLPT_SET((k_voronoi<false>));
    RLPT_SET((k_closest_hit<true, true>)); RLPT_SET((k_closest_hit<true, false>)); RLPT_SET((k_closest_hit<false, true>)); RLPT_SET((k_closest_hit<false, false>));
    RLPT_SET((k_closest_hit_bvh<true>)); RLPT_SET((k_closest_hit_bvh<false>));
#undef RLPT_SET
    return (int)e;
}

// ------------------------------------------------------------------------------------------------ Q merge + CDF rebuild
// One warp per radiance volume. Consumes the (all-reduced) accumulators: running-mean merge
//   Q <- (visits*Q + sum_targets) / (visits + count), clamp at RADIANCE_THRESHOLD, visits += count
// (the closed form of the reference's per-visit alpha = 1/(1+visits) update, radiance_volume.cu:283-301), then
// update_radiance_distribution (radiance_volume.cu:149-188) as a warp prefix sum, and the irradiance estimate
// sum_k Q_k cos_k luminance/pi (radiance_volume.cu:57-68) recomputed from cell centres.
__global__ void __launch_bounds__(BLOCK) k_merge_cdf(RadianceDev rm, const float* __restrict__ surf_lum_over_pi, float threshold, int rebuild_only) {
    const unsigned full = 0xffffffffu;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int vol = warp; vol < rm.n_vol; vol += nwarps) {
        size_t base = (size_t)vol * CELLS;
        float temp[5]; float qv[5]; float tsum = 0.f, irr = 0.f;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            int k = c * 32 + lane; temp[c] = 0.f; qv[c] = 0.f;
            if (k < CELLS) {
                float q = rm.q[base + k];
                if (!rebuild_only) {
                    uint32_t cnt = rm.acc_cnt[base + k];
                    if (cnt) {
                        float vs = (float)rm.visits[base + k];
                        q = (vs * q + rm.acc_sum[base + k]) / (vs + (float)cnt);
                        q = q > threshold ? q : threshold;
                        rm.q[base + k] = q; rm.visits[base + k] += cnt;
                        rm.acc_cnt[base + k] = 0u; rm.acc_sum[base + k] = 0.f;
                    }
                }
                float w = q * c_cell_cos[k];
                irr += w;
                temp[c] = w > 0.f ? w : 0.f;                 // DISTRIBUTION_THRESHOLD = 0
                tsum += temp[c];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { tsum += __shfl_xor_sync(full, tsum, o); irr += __shfl_xor_sync(full, irr, o); }
        float total = 0.0000000001f + tsum;
        float carry = 0.f;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            float x = temp[c] / total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { float y = __shfl_up_sync(full, x, o); if (lane >= o) x += y; }
            x += carry;
            int k = c * 32 + lane;
            if (k < CELLS) { rm.cdf[base + k] = x; if (k % GRID == GRID - 1) rm.cdf_rows[(size_t)vol * GRID + k / GRID] = x; }
            carry = __shfl_sync(full, x, 31);
        }
        if (lane == 0) rm.irradiance[vol] = irr * __ldg(surf_lum_over_pi + __ldg(rm.vol_surface + vol));
    }
}
void launch_merge(const RadianceDev& rm, const float* surf_lum_over_pi, float threshold, int rebuild_only, cudaStream_t s) {
    int warps_per_block = BLOCK / 32;
    int grid = (rm.n_vol + warps_per_block - 1) / warps_per_block; if (grid > 148 * 16) grid = 148 * 16; if (grid < 1) grid = 1;
    k_merge_cdf<<<grid, BLOCK, 0, s>>>(rm, surf_lum_over_pi, threshold, rebuild_only);
}

// ------------------------------------------------------------------------------------------------ exchange + merge over peer memory
// Multi-GPU form of k_merge_cdf: instead of all-reducing the accumulators (1152 B per volume) with a collective library and
// merging afterwards, every rank reduces ITS slice of the volumes straight out of all ranks' accumulators (P2P loads over
// NVLink / NVSwitch), merges, rebuilds the CDFs of the slice and stores the results into every rank's tables (P2P stores):
// reduce-scatter, the merge and the all-gather in one kernel, 1/N of the all-reduce's traffic per GPU on the way in, only
// the volumes that were actually visited on the way out. Replicas end bit-identical by construction (one owner computes
// each volume; the sum runs over the ranks in rank order). Each rank clears its own accumulators after the exchange. Synchronisation is two flag rounds in peer memory:
//   start  block 0 announces "my accumulators of this epoch are complete" (the kernel is stream-ordered after the tracing)
//          to every rank; every CTA waits until all ranks have announced
//   end    the last CTA to finish announces "my slice is stored everywhere"; k_wait_peers (next in the stream) waits for
//          all ranks' announcements, after which the local tables are complete and nobody reads the local accumulators
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
// A rank that never arrives (it skipped its merge, failed before it, or rebuilt its map on its own) must not hang the others inside a
// kernel: every wait gives up after P2P_TIMEOUT_NS of %globaltimer and raises the rank's error word (pt.error, checked by the
// host after every merge; the exchange is then marked broken and the frame fails with RLPT_ERR_COLLECTIVE).
#ifndef RLPT_P2P_REMOTE_ZERO
#define RLPT_P2P_REMOTE_ZERO 0        // 1: the consuming rank clears the accumulator cells through peer stores (round 1's form)
#endif
constexpr unsigned long long P2P_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool wait_flag(const unsigned* p, unsigned epoch, unsigned* error) {
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(p) - epoch) < 0) {
        __nanosleep(200);
        if (globaltimer_ns() - t0 > P2P_TIMEOUT_NS || *reinterpret_cast<volatile unsigned*>(error) != 0u) { atomicExch(error, 1u); return false; }
    }
    return true;
}
__global__ void __launch_bounds__(BLOCK) k_merge_cdf_p2p(RadianceDev rm, const __grid_constant__ PeerTables pt, const float* __restrict__ surf_lum_over_pi, float threshold,
                                                         unsigned epoch, unsigned* done_counter) {
    const unsigned full = 0xffffffffu;
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) { __threadfence_system(); for (int r = 0; r < pt.world; ++r) st_release_sys(pt.flags[r] + pt.rank, epoch); }
        const unsigned* mine = pt.flags[pt.rank];
        bool ok = true;
        for (int r = 0; r < pt.world && ok; ++r) ok = wait_flag(mine + r, epoch, pt.error);
        s_ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) return;                                                    // a peer never announced: touch nothing of anybody's tables
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int v0 = (int)((long long)rm.n_vol * pt.rank / pt.world), v1 = (int)((long long)rm.n_vol * (pt.rank + 1) / pt.world);
    for (int vol = v0 + warp; vol < v1; vol += nwarps) {
        const size_t base = (size_t)vol * CELLS;
        float temp[5]; float tsum = 0.f, irr = 0.f; bool touched = false;
        // all remote loads of the volume are issued before any is used: one NVLink round trip per volume, not one per cell and rank
        uint32_t cn[5][MAX_PEERS]; float sm[5][MAX_PEERS];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int k = c * 32 + lane;
#pragma unroll
            for (int r = 0; r < MAX_PEERS; ++r) {
                const bool on = r < pt.world && k < CELLS;
                cn[c][r] = on ? __ldcg(pt.acc_cnt[r] + base + k) : 0u;
                sm[c][r] = on ? __ldcg(pt.acc_sum[r] + base + k) : 0.f;
            }
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int k = c * 32 + lane; temp[c] = 0.f;
            if (k < CELLS) {
                float q = rm.q[base + k];
                uint32_t cnt = 0; float sum = 0.f;
#pragma unroll
                for (int r = 0; r < MAX_PEERS; ++r) {                     // rank order: the same sum on whichever rank owns the volume
#if RLPT_P2P_REMOTE_ZERO
                    if (r < pt.world && cn[c][r]) { cnt += cn[c][r]; sum += sm[c][r]; pt.acc_cnt[r][base + k] = 0u; pt.acc_sum[r][base + k] = 0.f; }
#else
                    if (r < pt.world && cn[c][r]) { cnt += cn[c][r]; sum += sm[c][r]; }       // (the accumulators are cleared by their owner afterwards: launch_merge_p2p)
#endif
                }
                if (cnt) {
                    const float vs = (float)rm.visits[base + k];
                    q = (vs * q + sum) / (vs + (float)cnt);
                    q = q > threshold ? q : threshold;
                    const uint32_t nvis = rm.visits[base + k] + cnt;
                    for (int r = 0; r < pt.world; ++r) { pt.q[r][base + k] = q; pt.visits[r][base + k] = nvis; }
                    touched = true;
                }
                const float w = q * c_cell_cos[k];
                irr += w;
                temp[c] = w > 0.f ? w : 0.f;
                tsum += temp[c];
            }
        }
        if (!__any_sync(full, touched)) continue;                        // nothing of this volume was visited: every rank's CDF stands
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { tsum += __shfl_xor_sync(full, tsum, o); irr += __shfl_xor_sync(full, irr, o); }
        const float total = 0.0000000001f + tsum;
        float carry = 0.f;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            float x = temp[c] / total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { float y = __shfl_up_sync(full, x, o); if (lane >= o) x += y; }
            x += carry;
            const int k = c * 32 + lane;
            if (k < CELLS) {
                for (int r = 0; r < pt.world; ++r) { pt.cdf[r][base + k] = x; if (k % GRID == GRID - 1) pt.cdf_rows[r][(size_t)vol * GRID + k / GRID] = x; }
            }
            carry = __shfl_sync(full, x, 31);
        }
        if (lane == 0) { const float e = irr * __ldg(surf_lum_over_pi + __ldg(rm.vol_surface + vol)); for (int r = 0; r < pt.world; ++r) pt.irradiance[r][vol] = e; }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {
            *done_counter = 0u; __threadfence_system();
            for (int r = 0; r < pt.world; ++r) st_release_sys(pt.flags[r] + MAX_PEERS + pt.rank, epoch);
        }
    }
}
__global__ void k_wait_peers(const unsigned* __restrict__ flags, int world, unsigned epoch, unsigned* error) {
    if (threadIdx.x < (unsigned)world) wait_flag(flags + MAX_PEERS + threadIdx.x, epoch, error);
}
void launch_merge_p2p(const RadianceDev& rm, const PeerTables& pt, const float* surf_lum_over_pi, float threshold, unsigned epoch, unsigned* done_counter, cudaStream_t s) {
    const int slice = rm.n_vol / pt.world + 1, warps_per_block = BLOCK / 32;
    int grid = (slice + warps_per_block - 1) / warps_per_block; if (grid > 148 * 8) grid = 148 * 8; if (grid < 1) grid = 1;
    k_merge_cdf_p2p<<<grid, BLOCK, 0, s>>>(rm, pt, surf_lum_over_pi, threshold, epoch, done_counter);
    k_wait_peers<<<1, 32, 0, s>>>(pt.flags[pt.rank], pt.world, epoch, pt.error);
    // Every rank has now read this rank's accumulators (its "slice stored" announcement comes after its loads): they are cleared HERE, locally (2 x 14 MB of
    // HBM writes for Cornell), instead of cell by cell through peer stores by whichever rank consumed them -- at 8 GPUs that was 7/8 of 28 MB of NVLink stores
    // per rank and frame inside the exchange kernel.
#if !RLPT_P2P_REMOTE_ZERO
    cudaMemsetAsync(rm.acc_sum, 0, sizeof(float) * (size_t)rm.n_vol * CELLS, s);
    cudaMemsetAsync(rm.acc_cnt, 0, sizeof(uint32_t) * (size_t)rm.n_vol * CELLS, s);
#endif
}

// ------------------------------------------------------------------------------------------------ frame buffer
__global__ void k_frame_mean(const float4* __restrict__ accum, float* __restrict__ rgb, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 a = accum[i]; float inv = a.w > 0.f ? 1.f / a.w : 0.f; rgb[3 * i] = a.x * inv; rgb[3 * i + 1] = a.y * inv; rgb[3 * i + 2] = a.z * inv; }
}
void launch_frame_mean(const float4* accum, float* rgb, int n, cudaStream_t s) { k_frame_mean<<<(n + 255) / 256, 256, 0, s>>>(accum, rgb, n); }

// SDLScreen::PutPixelSDL (G/sdl/sdl_screen.cpp:96-108): (128<<24) + (r<<16) + (g<<8) + b with r = uint32(clamp(255*c, 0, 255)),
// written at buffer[y*width + x] from pixel x*height + y (G/main.cu:235-239).
__global__ void k_pack_argb(const float4* __restrict__ accum, uint32_t* __restrict__ argb, int width, int height) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= width * height) return;
    int y = o / width, x = o % width;
    float4 a = accum[x * height + y]; float inv = a.w > 0.f ? 1.f / a.w : 0.f;
    uint32_t r = (uint32_t)fminf(fmaxf(255.f * (a.x * inv), 0.f), 255.f), g = (uint32_t)fminf(fmaxf(255.f * (a.y * inv), 0.f), 255.f), b = (uint32_t)fminf(fmaxf(255.f * (a.z * inv), 0.f), 255.f);
    argb[o] = (128u << 24) + (r << 16) + (g << 8) + b;
}
void launch_pack_argb(const float4* accum, uint32_t* argb, int width, int height, cudaStream_t s) {
    k_pack_argb<<<(width * height + 255) / 256, 256, 0, s>>>(accum, argb, width, height);
}

// ------------------------------------------------------------------------------------------------ FP32 peak probe
// 8 independent FMA chains per thread, 2 flop per FMA: the FP32-pipe roofline denominator measured on this GPU at the
// clocks it actually holds (MEASURED_PEAKS.json has no FP32 figure).
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
void launch_fp32_peak(float* out, int iters, int grid, cudaStream_t s) { k_fp32_peak<<<grid, 256, 0, s>>>(out, iters); }

}  // namespace rlpt
