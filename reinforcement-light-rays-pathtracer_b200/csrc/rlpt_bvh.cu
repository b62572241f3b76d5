// rlpt_bvh.cu -- BVH construction on the GPU (sm_100a) for the SoA triangle buffer the scene upload emits.
// The reference has no acceleration structure (brute force over every Surface and AreaLight, G/rays/ray.cu:22-35);
// this is new work asked for by BASELINE.json north_star.
//
// Two builders, both on the GPU, both emitting the same traversal layout:
//   * top-down binned SAH by one CTA (k_sah_build, below) for scenes up to SAH_MAX_PRIMS primitives -- the trees the tracing
//     kernels actually walk for the bundled scenes;
//   * linear BVH (Karras 2012) above that: 30-bit Morton code of each primitive's centroid, radix sort (CUB), one thread per
//     internal node finds its range and split from the sorted keys, AABBs fitted bottom-up with one atomic flag per node.
// Nodes are emitted breadth-first:
//   node = 4 x float4:  (c0.min.xyz, c0.max.x) (c0.max.yz, c1.min.xy) (c1.min.z, c1.max.xyz) (as_float(c0), as_float(c1), 0, 0)
//   child link >= 0: node index;  < 0: ~primitive id (one primitive per leaf)
// Leaf boxes are padded (kPad * primitive extent + kAbs) so that traversal can never cull a primitive that the
// reference's FP32 Cramer test would accept for a ray passing just outside the exact triangle.
//
// The kernels do not walk that binary tree: k_collapse4 (below) folds it, still on the GPU, into a 4-WIDE tree -- a wide node
// absorbs the binary nodes under it (largest surface area first) until it has four children; a binary subtree of at most
// `leaf_max` primitives becomes one leaf. Half the dependent node visits per ray, and the per-visit bookkeeping (stack, ordering,
// loop control) is paid once per four boxes. Wide node = 7 x float4:
//   [0] lo.x of children 0..3   [1] hi.x   [2] lo.y   [3] hi.y   [4] lo.z   [5] hi.z        (a ray picks near / far by its octant)
//   [6] links: inner children FIRST, and their nodes are consecutive (link_k = link_0 + k), so the traversal stack and the
//       ordering keys carry a slot number instead of a link; leaf: ~(first record << 3 | records), the records being consecutive
//       in tri4 (the triangle buffer re-ordered by leaf, primitive id in the third float4's z); unused slot: BVH4_EMPTY (= a leaf of zero
//       records) + an inverted box (lo = FLT_MAX, hi = -FLT_MAX) that no finite ray can enter.
#include "rlpt_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <vector>
#include <queue>
#include <cfloat>
#include <cstring>
#include <cstdlib>

namespace rlpt {

namespace {
constexpr float kPad = 1.f / 128.f;
constexpr float kAbs = 1e-4f;

struct Box { float lo[3], hi[3]; };

__global__ void k_prim_bounds(const float4* __restrict__ tri, int n, Box* __restrict__ boxes, float* __restrict__ scene /*6*/) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = tri[3 * i], b = tri[3 * i + 1], c = tri[3 * i + 2];
    float v0[3] = { a.x, a.y, a.z }, e1[3] = { a.w, b.x, b.y }, e2[3] = { b.z, b.w, c.x };
    Box bx; float ext = 0.f;
    for (int k = 0; k < 3; ++k) {
        float p1 = v0[k] + e1[k], p2 = v0[k] + e2[k];
        bx.lo[k] = fminf(v0[k], fminf(p1, p2)); bx.hi[k] = fmaxf(v0[k], fmaxf(p1, p2));
        ext = fmaxf(ext, bx.hi[k] - bx.lo[k]);
    }
    float pad = kPad * ext + kAbs;
    for (int k = 0; k < 3; ++k) { bx.lo[k] -= pad + 1e-6f * fabsf(bx.lo[k]); bx.hi[k] += pad + 1e-6f * fabsf(bx.hi[k]); }
    boxes[i] = bx;
    // scene bounds: float atomics via ordered-int trick are not needed for <= thousands of primitives; use CAS min/max
    for (int k = 0; k < 3; ++k) {
        int* lo = reinterpret_cast<int*>(scene + k); int* hi = reinterpret_cast<int*>(scene + 3 + k);
        int old = *lo; while (__int_as_float(old) > bx.lo[k]) { int prev = atomicCAS(lo, old, __float_as_int(bx.lo[k])); if (prev == old) break; old = prev; }
        old = *hi; while (__int_as_float(old) < bx.hi[k]) { int prev = atomicCAS(hi, old, __float_as_int(bx.hi[k])); if (prev == old) break; old = prev; }
    }
}

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void k_morton(const Box* __restrict__ boxes, int n, const float* __restrict__ scene, uint64_t* __restrict__ keys, int* __restrict__ ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t code = 0;
    for (int k = 0; k < 3; ++k) {
        float lo = scene[k], hi = scene[3 + k], c = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
        float u = hi > lo ? (c - lo) / (hi - lo) : 0.5f;
        uint32_t q = (uint32_t)fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
        code |= expand_bits(q) << (2 - k);
    }
    keys[i] = ((uint64_t)code << 32) | (uint32_t)i;        // primitive id in the low word makes every key unique
    ids[i] = i;
}

__device__ __forceinline__ int delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);
}
// Karras 2012, one thread per internal node i in [0, n-1): children/parents in the temporary "LBVH numbering":
// internal nodes 0..n-2, leaves encoded as (n-1) + sorted position.
__global__ void k_karras(const uint64_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right, int* __restrict__ parent, int* __restrict__ rng_b, int* __restrict__ rng_e) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2; while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0; for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0; int t = l;
    do { t = (t + 1) >> 1; if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? (n - 1) + gamma : gamma;
    int rc = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
    left[i] = lc; right[i] = rc; parent[lc] = i; parent[rc] = i;
    rng_b[i] = lo; rng_e[i] = hi + 1;                        // the node's primitives: sorted positions [lo, hi]
    if (i == 0) parent[0] = -1;
}
// bottom-up AABB fit: each leaf walks up; the second thread to reach a node merges its children's boxes
__global__ void k_fit(const Box* __restrict__ prim_boxes, const int* __restrict__ ids, int n, const int* __restrict__ left, const int* __restrict__ right,
                      const int* __restrict__ parent, Box* __restrict__ node_boxes /* 2n-1 */, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = (n - 1) + i;
    node_boxes[node] = prim_boxes[ids[i]];
    __threadfence();
    int p = parent[node];
    while (p >= 0) {
        if (atomicAdd(flags + p, 1) == 0) return;
        Box a = node_boxes[left[p]], b = node_boxes[right[p]], m;
        for (int k = 0; k < 3; ++k) { m.lo[k] = fminf(a.lo[k], b.lo[k]); m.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
        node_boxes[p] = m;
        __threadfence();
        p = parent[p];
    }
}
// ---- top-down binned SAH build, one CTA (scenes up to SAH_MAX_PRIMS primitives; the LBVH above serves the rest).
// The Morton-order tree costs a ray of Medieval_House 23 node visits and 7 triangle solves on average, and single rays
// several hundred visits -- the dependent chain that sets the floor of every small launch. The surface-area heuristic
// (16 centroid bins per axis, all three axes, one primitive per leaf as before) is the classic cure. Nodes are processed
// breadth-first by the whole CTA: bounds -> binning with shared-memory atomics -> 45 candidate planes costed in parallel ->
// stable partition of the index range with a block scan. Everything that decides a split is a set operation (min, max,
// count), so the tree does not depend on thread scheduling. It writes the LBVH's intermediate form (left/right links with
// leaves as (n-1) + position, node boxes), so numbering and k_emit are shared.
constexpr int SAH_T = 1024, SAH_BINS = 16, SAH_MAX_DEPTH = 29, SAH_MAX_PRIMS = 1 << 17;
struct SahItem { int b, e, depth; };
__device__ __forceinline__ int f2o(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }       // order-preserving float -> int
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ int sah_bin(const Box& bx, int k, float cmin, float scale) {
    const float c = 0.5f * (bx.lo[k] + bx.hi[k]);
    return scale > 0.f ? min(SAH_BINS - 1, max(0, (int)((c - cmin) * scale))) : 0;
}
__device__ __forceinline__ int levels_for(int c) { return c <= 1 ? 0 : 32 - __clz(c - 1); }                        // internal levels a median build of c primitives needs
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = s_warp[lane];
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
        s_warp[lane] = t;
    }
    __syncthreads();
    const int pre = (w > 0 ? s_warp[w - 1] : 0) + x - v;
    total = s_warp[31];
    __syncthreads();
    return pre;
}
__global__ void __launch_bounds__(SAH_T) k_sah_build(const Box* __restrict__ boxes, int n, int* __restrict__ ids, int* __restrict__ tmp, SahItem* __restrict__ queue,
                                                     int* __restrict__ left, int* __restrict__ right, Box* __restrict__ node_boxes, int* __restrict__ depth_out, int* __restrict__ rng_b, int* __restrict__ rng_e) {
    __shared__ int s_cb[6], s_nb[6];
    __shared__ int s_cnt[3][SAH_BINS], s_lo[3][SAH_BINS][3], s_hi[3][SAH_BINS][3];
    __shared__ float s_cost[48];
    __shared__ int s_axis, s_split, s_nl, s_next, s_maxd, s_warp[32];
    const int tid = threadIdx.x;
    for (int j = tid; j < n; j += SAH_T) ids[j] = j;
    if (tid == 0) { queue[0] = SahItem{ 0, n, 1 }; s_next = 1; s_maxd = 1; }
    __syncthreads();
    for (int node = 0; node < n - 1; ++node) {
        const SahItem it = queue[node];
        const int b = it.b, e = it.e, c = e - b;
        if (tid == 0) { rng_b[node] = b; rng_e[node] = e; }
        if (tid < 6) { s_cb[tid] = tid < 3 ? 0x7fffffff : (int)0x80000000; s_nb[tid] = s_cb[tid]; }
        for (int k = tid; k < 3 * SAH_BINS; k += SAH_T) {
            (&s_cnt[0][0])[k] = 0;
            for (int d = 0; d < 3; ++d) { (&s_lo[0][0][0])[3 * k + d] = 0x7fffffff; (&s_hi[0][0][0])[3 * k + d] = (int)0x80000000; }
        }
        __syncthreads();
        for (int j = b + tid; j < e; j += SAH_T) {
            const Box bx = boxes[ids[j]];
            for (int k = 0; k < 3; ++k) {
                const int cc = f2o(0.5f * (bx.lo[k] + bx.hi[k]));
                atomicMin(&s_cb[k], cc); atomicMax(&s_cb[3 + k], cc);
                atomicMin(&s_nb[k], f2o(bx.lo[k])); atomicMax(&s_nb[3 + k], f2o(bx.hi[k]));
            }
        }
        __syncthreads();
        float cmin[3], scale[3];
        for (int k = 0; k < 3; ++k) {
            const float lo = o2f(s_cb[k]), hi = o2f(s_cb[3 + k]);
            cmin[k] = lo; scale[k] = hi > lo ? (float)SAH_BINS / (hi - lo) : 0.f;
        }
        if (tid == 0) { Box m; for (int k = 0; k < 3; ++k) { m.lo[k] = o2f(s_nb[k]); m.hi[k] = o2f(s_nb[3 + k]); } node_boxes[node] = m; }
        for (int j = b + tid; j < e; j += SAH_T) {
            const Box bx = boxes[ids[j]];
            for (int k = 0; k < 3; ++k) {
                const int bin = sah_bin(bx, k, cmin[k], scale[k]);
                atomicAdd(&s_cnt[k][bin], 1);
                for (int d = 0; d < 3; ++d) { atomicMin(&s_lo[k][bin][d], f2o(bx.lo[d])); atomicMax(&s_hi[k][bin][d], f2o(bx.hi[d])); }
            }
        }
        __syncthreads();
        if (tid < 3 * (SAH_BINS - 1)) {
            const int k = tid / (SAH_BINS - 1), sp = tid % (SAH_BINS - 1);          // left = bins 0..sp, right = sp+1..
            float cost = FLT_MAX;
            if (scale[k] > 0.f) {
                int nl = 0, nr = 0; float llo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, lhi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX }, rlo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, rhi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
                for (int q = 0; q < SAH_BINS; ++q) {
                    const int cnt = s_cnt[k][q]; if (!cnt) continue;
                    if (q <= sp) { nl += cnt; for (int d = 0; d < 3; ++d) { llo[d] = fminf(llo[d], o2f(s_lo[k][q][d])); lhi[d] = fmaxf(lhi[d], o2f(s_hi[k][q][d])); } }
                    else { nr += cnt; for (int d = 0; d < 3; ++d) { rlo[d] = fminf(rlo[d], o2f(s_lo[k][q][d])); rhi[d] = fmaxf(rhi[d], o2f(s_hi[k][q][d])); } }
                }
                if (nl > 0 && nr > 0) {
                    const float lx = lhi[0] - llo[0], ly = lhi[1] - llo[1], lz = lhi[2] - llo[2], rx = rhi[0] - rlo[0], ry = rhi[1] - rlo[1], rz = rhi[2] - rlo[2];
                    cost = (lx * ly + ly * lz + lz * lx) * (float)nl + (rx * ry + ry * rz + rz * rx) * (float)nr;
                }
            }
            s_cost[tid] = cost;
        }
        __syncthreads();
        if (tid == 0) {
            int best = -1; float bc = FLT_MAX;
            if (it.depth + levels_for(c - 1) <= SAH_MAX_DEPTH)                       // otherwise halve: the stack of the traversal holds 32 entries
                for (int q = 0; q < 3 * (SAH_BINS - 1); ++q) if (s_cost[q] < bc) { bc = s_cost[q]; best = q; }
            if (best >= 0) {
                s_axis = best / (SAH_BINS - 1); s_split = best % (SAH_BINS - 1);
                int nl = 0; for (int q = 0; q <= s_split; ++q) nl += s_cnt[s_axis][q];
                s_nl = nl;
            } else { s_axis = -1; s_nl = c / 2; }
        }
        __syncthreads();
        const int axis = s_axis, split = s_split, nl = s_nl;
        if (axis >= 0) {
            int lpos = b, rpos = b + nl;
            for (int cs = b; cs < e; cs += SAH_T) {
                const int j = cs + tid; const bool valid = j < e;
                const int id = valid ? ids[j] : 0;
                const bool pl = valid && sah_bin(boxes[id], axis, cmin[axis], scale[axis]) <= split;
                int total; const int off = block_excl_scan(pl ? 1 : 0, s_warp, total);
                if (valid) { if (pl) tmp[lpos + off] = id; else tmp[rpos + (tid - off)] = id; }
                lpos += total; rpos += min(SAH_T, e - cs) - total;
            }
            __syncthreads();
            for (int j = b + tid; j < e; j += SAH_T) ids[j] = tmp[j];
        }
        if (tid == 0) {
            const int m = b + nl;
            int links[2];
            for (int side = 0; side < 2; ++side) {
                const int lb = side ? m : b, le = side ? e : m;
                if (le - lb == 1) links[side] = (n - 1) + lb;
                else { const int id = s_next++; queue[id] = SahItem{ lb, le, it.depth + 1 }; links[side] = id; if (it.depth + 1 > s_maxd) s_maxd = it.depth + 1; }
            }
            left[node] = links[0]; right[node] = links[1];
        }
        __syncthreads();
    }
    for (int j = tid; j < n; j += SAH_T) node_boxes[(n - 1) + j] = boxes[ids[j]];
    if (tid == 0) *depth_out = s_maxd;
}

// emit traversal nodes; `order[k]` = LBVH internal node placed at output slot k (breadth-first), `slot_of[i]` its inverse
__global__ void k_emit(const Box* __restrict__ node_boxes, const int* __restrict__ ids, int n, const int* __restrict__ left, const int* __restrict__ right,
                       const int* __restrict__ order, const int* __restrict__ slot_of, float4* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n - 1) return;
    int i = order[k];
    int lc = left[i], rc = right[i];
    Box a = node_boxes[lc], b = node_boxes[rc];
    int l0 = lc >= n - 1 ? ~ids[lc - (n - 1)] : slot_of[lc];
    int l1 = rc >= n - 1 ? ~ids[rc - (n - 1)] : slot_of[rc];
    out[4 * k + 0] = make_float4(a.lo[0], a.lo[1], a.lo[2], a.hi[0]);
    out[4 * k + 1] = make_float4(a.hi[1], a.hi[2], b.lo[0], b.lo[1]);
    out[4 * k + 2] = make_float4(b.lo[2], b.hi[0], b.hi[1], b.hi[2]);
    out[4 * k + 3] = make_float4(__int_as_float(l0), __int_as_float(l1), 0.f, 0.f);
}

// ---- binary -> 4-wide collapse, one CTA, level by level (the wide tree is numbered breadth-first: the children of one node get
// consecutive slots, which the traversal relies on). Task = (binary node, wide slot). Everything that decides the topology is a
// function of the binary tree alone; slots are handed out by a block scan, so the wide tree is the same on every run and GPU.
constexpr int C4_T = 1024;
struct C4Task { int bnode, wslot; };
__device__ __forceinline__ float box_area(const Box& b) {
    const float x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
    return x * y + y * z + z * x;
}
__global__ void __launch_bounds__(C4_T) k_collapse4(const Box* __restrict__ node_boxes, int n, const int* __restrict__ left, const int* __restrict__ right,
                                                    const int* __restrict__ rng_b, const int* __restrict__ rng_e, int leaf_max, C4Task* __restrict__ queue,
                                                    float4* __restrict__ out, int* __restrict__ counts /* [0] nodes, [1] depth */) {
    __shared__ int s_warp[32];
    __shared__ int s_begin, s_end, s_alloc, s_depth;
    const int tid = threadIdx.x;
    if (tid == 0) { queue[0] = C4Task{ 0, 0 }; s_begin = 0; s_end = 1; s_alloc = 1; s_depth = 0; }
    __syncthreads();
    auto is_inner = [&](int c) { return c < n - 1 && rng_e[c] - rng_b[c] > leaf_max; };       // becomes a wide node of its own
    while (true) {
        const int begin = s_begin, end = s_end;
        if (begin >= end) break;
        int qtail = end;
        for (int chunk = begin; chunk < end; chunk += C4_T) {
            const int t = chunk + tid; const bool valid = t < end;
            int c[4] = { -1, -1, -1, -1 }; int nc = 0, n_inner = 0; C4Task task{ 0, 0 };
            if (valid) {
                task = queue[t];
                c[0] = left[task.bnode]; c[1] = right[task.bnode]; nc = 2;
                while (nc < 4) {                                                          // open the inner child with the largest box
                    int pick = -1; float best = -1.f;
                    for (int k = 0; k < nc; ++k) if (is_inner(c[k])) { const float a = box_area(node_boxes[c[k]]); if (a > best) { best = a; pick = k; } }
                    if (pick < 0) break;
                    const int o = c[pick]; c[pick] = left[o]; c[nc++] = right[o];
                }
                // inner children first (stable), then leaves
                int ord[4], m = 0;
                for (int k = 0; k < nc; ++k) if (is_inner(c[k])) ord[m++] = c[k];
                n_inner = m;
                for (int k = 0; k < nc; ++k) if (!is_inner(c[k])) ord[m++] = c[k];
                for (int k = 0; k < nc; ++k) c[k] = ord[k];
            }
            int total; const int off = block_excl_scan(n_inner, s_warp, total);
            const int base = s_alloc + off, qpos = qtail + off;
            if (valid) {
                float pl[6][4]; int link[4];
                for (int k = 0; k < 4; ++k) {
                    if (k < nc) {
                        const Box b = node_boxes[c[k]];
                        for (int d = 0; d < 3; ++d) { pl[2 * d][k] = b.lo[d]; pl[2 * d + 1][k] = b.hi[d]; }
                        if (k < n_inner) { link[k] = base + k; queue[qpos + k] = C4Task{ c[k], base + k }; }
                        else if (c[k] >= n - 1) link[k] = ~(((c[k] - (n - 1)) << 3) | 1);                       // one primitive: its sorted position
                        else link[k] = ~((rng_b[c[k]] << 3) | (rng_e[c[k]] - rng_b[c[k]]));                     // a small subtree: its run of positions
                    } else {
                        for (int d = 0; d < 3; ++d) { pl[2 * d][k] = FLT_MAX; pl[2 * d + 1][k] = -FLT_MAX; }
                        link[k] = BVH4_EMPTY;
                    }
                }
                float4* o = out + 7 * (size_t)task.wslot;
                for (int r = 0; r < 6; ++r) o[r] = make_float4(pl[r][0], pl[r][1], pl[r][2], pl[r][3]);
                o[6] = make_float4(__int_as_float(link[0]), __int_as_float(link[1]), __int_as_float(link[2]), __int_as_float(link[3]));
            }
            __syncthreads();
            if (tid == 0) s_alloc += total;
            qtail += total;
            __syncthreads();
        }
        if (tid == 0) { s_begin = end; s_end = qtail; s_depth++; }
        __syncthreads();
    }
    if (tid == 0) { counts[0] = s_alloc; counts[1] = s_depth; }
}
// triangle records in leaf order, primitive id in the spare word
__global__ void k_tri4(const float4* __restrict__ tri, const int* __restrict__ ids, int n, float4* __restrict__ tri4) {
    int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const int gid = ids[pos];
    float4 c = tri[3 * gid + 2]; c.z = __int_as_float(gid);
    tri4[3 * pos] = tri[3 * gid]; tri4[3 * pos + 1] = tri[3 * gid + 1]; tri4[3 * pos + 2] = c;
}
}  // namespace

#define BVH_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = (int)e_; goto done; } } while (0)

int bvh_build_gpu(const float4* d_tri, int n, int leaf_max, float4** d_bvh, int* n_nodes, int* depth, float4** d_bvh4, int* n_nodes4, int* depth4, float4** d_tri4, cudaStream_t s) {
    int rc = 0;
    Box *boxes = nullptr, *node_boxes = nullptr; float* scene = nullptr; uint64_t *keys = nullptr, *keys_s = nullptr; int *ids = nullptr, *ids_s = nullptr;
    int *left = nullptr, *right = nullptr, *parent = nullptr, *flags = nullptr, *order = nullptr, *slot_of = nullptr; void* tmp = nullptr; size_t tmp_bytes = 0;
    int *rng_b = nullptr, *rng_e = nullptr, *c4_counts = nullptr; C4Task* c4_queue = nullptr; float4 *out4 = nullptr, *tri4 = nullptr;
    float4* out = nullptr; SahItem* sah_queue = nullptr; int* sah_depth = nullptr; bool use_sah = false;
    const int T = 128;
    *d_bvh = nullptr; *n_nodes = 0; *depth = 0; *d_bvh4 = nullptr; *n_nodes4 = 0; *depth4 = 0; *d_tri4 = nullptr;
    if (n <= 0) return 0;
    leaf_max = leaf_max < 1 ? 1 : (leaf_max > BVH4_LEAF_MAX ? BVH4_LEAF_MAX : leaf_max);
    BVH_CK(cudaMalloc(&boxes, sizeof(Box) * n)); BVH_CK(cudaMalloc(&scene, sizeof(float) * 6));
    {
        float init[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
        BVH_CK(cudaMemcpyAsync(scene, init, sizeof init, cudaMemcpyHostToDevice, s));
    }
    k_prim_bounds<<<(n + T - 1) / T, T, 0, s>>>(d_tri, n, boxes, scene);
    BVH_CK(cudaMalloc(&tri4, sizeof(float4) * 3 * (size_t)n));
    if (n == 1) {
        // a single primitive: one node whose second child is an empty box
        Box b; BVH_CK(cudaStreamSynchronize(s)); BVH_CK(cudaMemcpy(&b, boxes, sizeof(Box), cudaMemcpyDeviceToHost));
        float4 h[4] = { make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]), make_float4(b.hi[1], b.hi[2], FLT_MAX, FLT_MAX),
                        make_float4(FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX), make_float4(0, 0, 0, 0) };
        int l0 = ~0; memcpy(&h[3].x, &l0, 4); memcpy(&h[3].y, &l0, 4);
        BVH_CK(cudaMalloc(&out, sizeof(float4) * 4)); BVH_CK(cudaMemcpy(out, h, sizeof h, cudaMemcpyHostToDevice));
        float4 w[7]; const int leaf = ~((0 << 3) | 1), none = BVH4_EMPTY;
        for (int d = 0; d < 3; ++d) { w[2 * d] = make_float4(b.lo[d], FLT_MAX, FLT_MAX, FLT_MAX); w[2 * d + 1] = make_float4(b.hi[d], -FLT_MAX, -FLT_MAX, -FLT_MAX); }
        memcpy(&w[6].x, &leaf, 4); memcpy(&w[6].y, &none, 4); memcpy(&w[6].z, &none, 4); memcpy(&w[6].w, &none, 4);
        BVH_CK(cudaMalloc(&out4, sizeof(float4) * 7)); BVH_CK(cudaMemcpy(out4, w, sizeof w, cudaMemcpyHostToDevice));
        BVH_CK(cudaMalloc(&ids, sizeof(int))); BVH_CK(cudaMemset(ids, 0, sizeof(int)));
        k_tri4<<<1, T, 0, s>>>(d_tri, ids, 1, tri4);
        BVH_CK(cudaStreamSynchronize(s)); BVH_CK(cudaGetLastError());
        *d_bvh = out; out = nullptr; *n_nodes = 1; *depth = 1; *d_bvh4 = out4; out4 = nullptr; *n_nodes4 = 1; *depth4 = 1; *d_tri4 = tri4; tri4 = nullptr; goto done;
    }
    BVH_CK(cudaMalloc(&ids, sizeof(int) * n)); BVH_CK(cudaMalloc(&ids_s, sizeof(int) * n));
    BVH_CK(cudaMalloc(&left, sizeof(int) * (n - 1))); BVH_CK(cudaMalloc(&right, sizeof(int) * (n - 1)));
    BVH_CK(cudaMalloc(&rng_b, sizeof(int) * (n - 1))); BVH_CK(cudaMalloc(&rng_e, sizeof(int) * (n - 1)));
    BVH_CK(cudaMalloc(&node_boxes, sizeof(Box) * (2 * n - 1)));
    {
        const char* e = getenv("RLPT_BVH_BUILD");                    // "lbvh": the Morton-order build for every size (A/B runs)
        use_sah = n <= SAH_MAX_PRIMS && !(e && !strcmp(e, "lbvh"));
    }
    if (use_sah) {
        BVH_CK(cudaMalloc(&sah_queue, sizeof(SahItem) * (n - 1))); BVH_CK(cudaMalloc(&sah_depth, sizeof(int)));
        k_sah_build<<<1, SAH_T, 0, s>>>(boxes, n, ids_s, ids, sah_queue, left, right, node_boxes, sah_depth, rng_b, rng_e);
    } else {
        BVH_CK(cudaMalloc(&keys, sizeof(uint64_t) * n)); BVH_CK(cudaMalloc(&keys_s, sizeof(uint64_t) * n));
        k_morton<<<(n + T - 1) / T, T, 0, s>>>(boxes, n, scene, keys, ids);
        BVH_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_s, ids, ids_s, n, 0, 64, s));
        BVH_CK(cudaMalloc(&tmp, tmp_bytes));
        BVH_CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_s, ids, ids_s, n, 0, 64, s));
        BVH_CK(cudaMalloc(&parent, sizeof(int) * (2 * n - 1))); BVH_CK(cudaMalloc(&flags, sizeof(int) * (n - 1)));
        BVH_CK(cudaMemsetAsync(flags, 0, sizeof(int) * (n - 1), s));
        k_karras<<<(n - 1 + T - 1) / T, T, 0, s>>>(keys_s, n, left, right, parent, rng_b, rng_e);
        k_fit<<<(n + T - 1) / T, T, 0, s>>>(boxes, ids_s, n, left, right, parent, node_boxes, flags);
    }
    // the 4-wide tree the kernels walk, and the triangle records in its leaf order
    BVH_CK(cudaMalloc(&c4_queue, sizeof(C4Task) * (n - 1))); BVH_CK(cudaMalloc(&c4_counts, sizeof(int) * 2));
    BVH_CK(cudaMalloc(&out4, sizeof(float4) * 7 * (size_t)(n - 1)));
    k_collapse4<<<1, C4_T, 0, s>>>(node_boxes, n, left, right, rng_b, rng_e, leaf_max, c4_queue, out4, c4_counts);
    k_tri4<<<(n + T - 1) / T, T, 0, s>>>(d_tri, ids_s, n, tri4);
    {
        // breadth-first numbering of the internal nodes (host walk over the n-1 child links; O(n), one-off)
        std::vector<int> hl(n - 1), hr(n - 1), ord, slot(n - 1, -1);
        BVH_CK(cudaStreamSynchronize(s));
        BVH_CK(cudaMemcpy(hl.data(), left, sizeof(int) * (n - 1), cudaMemcpyDeviceToHost));
        BVH_CK(cudaMemcpy(hr.data(), right, sizeof(int) * (n - 1), cudaMemcpyDeviceToHost));
        std::queue<std::pair<int, int>> q; q.push({ 0, 1 }); int maxd = 1;
        while (!q.empty()) {
            auto [i, d] = q.front(); q.pop();
            slot[i] = (int)ord.size(); ord.push_back(i); if (d > maxd) maxd = d;
            if (hl[i] < n - 1) q.push({ hl[i], d + 1 });
            if (hr[i] < n - 1) q.push({ hr[i], d + 1 });
        }
        if ((int)ord.size() != n - 1) { rc = -2; goto done; }
        BVH_CK(cudaMalloc(&order, sizeof(int) * (n - 1))); BVH_CK(cudaMalloc(&slot_of, sizeof(int) * (n - 1)));
        BVH_CK(cudaMemcpy(order, ord.data(), sizeof(int) * (n - 1), cudaMemcpyHostToDevice));
        BVH_CK(cudaMemcpy(slot_of, slot.data(), sizeof(int) * (n - 1), cudaMemcpyHostToDevice));
        *depth = maxd + 1;
    }
    BVH_CK(cudaMalloc(&out, sizeof(float4) * 4 * (n - 1)));
    k_emit<<<(n - 1 + T - 1) / T, T, 0, s>>>(node_boxes, ids_s, n, left, right, order, slot_of, out);
    BVH_CK(cudaStreamSynchronize(s));
    BVH_CK(cudaGetLastError());
    {
        int hc[2] = { 0, 0 }; BVH_CK(cudaMemcpy(hc, c4_counts, sizeof hc, cudaMemcpyDeviceToHost));
        if (hc[0] < 1 || hc[0] > n - 1) { rc = -2; goto done; }
        *n_nodes4 = hc[0]; *depth4 = hc[1];
    }
    *d_bvh = out; out = nullptr; *n_nodes = n - 1; *d_bvh4 = out4; out4 = nullptr; *d_tri4 = tri4; tri4 = nullptr;
done:
    cudaFree(boxes); cudaFree(node_boxes); cudaFree(scene); cudaFree(keys); cudaFree(keys_s); cudaFree(ids); cudaFree(ids_s);
    cudaFree(left); cudaFree(right); cudaFree(parent); cudaFree(flags); cudaFree(order); cudaFree(slot_of); cudaFree(tmp); cudaFree(out); cudaFree(sah_queue); cudaFree(sah_depth);
    cudaFree(rng_b); cudaFree(rng_e); cudaFree(c4_counts); cudaFree(c4_queue); cudaFree(out4); cudaFree(tri4);
    return rc;
}

}  // namespace rlpt
