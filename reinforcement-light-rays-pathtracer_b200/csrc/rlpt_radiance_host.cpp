// rlpt_radiance_host.cpp -- host-side construction of the radiance map (one-off, outside the timed hot path):
// how many radiance volumes each surface gets, where they sit, and the kd-tree over them.
//
// These three steps are kept arithmetically identical to the reference so that the volumes, the flattened tree and
// therefore every nearest-volume answer are the reference's own, bit for bit:
//   RadianceMap::get_radiance_volumes_count          G/radiance_volumes/radiance_map.cu:60-67   floor(area / AREA_PER_SAMPLE)
//   Triangle::compute_area                           G/objects/triangle.cu:17-26
//   RadianceMap::uniformly_sample_radiance_volumes   radiance_map.cu:72-84 + Triangle::sample_position_on_plane triangle.cu:30-45
//                                                    (host rand(), never seeded => glibc's srand(1) stream)
//   RadianceTree::RadianceTree / convert_to_array    G/radiance_volumes/radiance_tree.cu:12-62,135-196 (std::sort per level)
// Built with -ffp-contract=off: one rounding per operator, as in the reference's -O0 host code.
#include "rlpt_radiance_host.h"
#include "rlpt_device.cuh"
#include <unordered_map>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace rlpt {

static inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

float host_triangle_area(const float* v /*9*/) {
    float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
    float len = std::sqrt(dot3(e1, e1)) * std::sqrt(dot3(e2, e2));
    float c = dot3(e1, e2) / len;
    float s = (float)std::sqrt(1 - std::pow((double)c, 2));     // pow(float,int) and sqrt run in double on the host
    return 0.5f * len * s;
}

void host_triangle_normal(const float* v, float* n) {            // G/objects/triangle.cu:67-76: normalize(cross(e2, e1))
    float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
    float c[3] = { e2[1] * e1[2] - e1[1] * e2[2], e2[2] * e1[0] - e1[2] * e2[0], e2[0] * e1[1] - e1[0] * e2[1] };
    float inv = 1.f / std::sqrt(dot3(c, c));
    n[0] = c[0] * inv; n[1] = c[1] * inv; n[2] = c[2] * inv;
}

namespace {
struct Builder {
    const std::vector<HostVolume>& vol;
    std::vector<HostTreeElement>& out;
    struct Node { int dim; float median; int volume; Node* l; Node* r; };
    std::vector<Node*> pool;
    ~Builder() { for (Node* n : pool) delete n; }
    float coord(int id, int dim) const { return vol[id].pos[dim]; }
    Node* build(std::vector<int>& ids, int dim) {
        Node* node = new Node{ dim, 0.f, -1, nullptr, nullptr }; pool.push_back(node);
        int n = (int)ids.size();
        if (n == 0) return node;
        if (n == 1) { node->median = coord(ids[0], dim); node->volume = ids[0]; return node; }
        std::sort(ids.begin(), ids.end(), [&](int a, int b) { return coord(a, dim) < coord(b, dim); });
        int mi;
        if (n % 2 == 0) { mi = n / 2 - 1; node->median = (coord(ids[mi], dim) + coord(ids[mi + 1], dim)) / 2; }
        else { mi = n / 2; node->median = coord(ids[mi], dim); }
        std::vector<int> L(ids.begin(), ids.begin() + mi + 1), R(ids.begin() + mi + 1, ids.end());
        node->l = build(L, (dim + 1) % 3); node->r = build(R, (dim + 1) % 3);
        return node;
    }
    void flatten(Node* t, int idx) {
        int last = (int)out.size() - 1;
        if (t->volume >= 0) {
            HostTreeElement e{}; e.dim = t->dim; e.leaf = 1; e.data = (float)t->volume;
            for (int k = 0; k < 3; ++k) { e.pos[k] = vol[t->volume].pos[k]; e.nrm[k] = vol[t->volume].nrm[k]; }
            out[idx] = e; return;
        }
        out[idx].left = (unsigned)(last + 1); out[idx].right = (unsigned)(last + 2);
        HostTreeElement a{}; a.dim = (t->dim + 1) % 3; a.data = t->l->median;
        HostTreeElement b = a; b.data = t->r->median;
        out.push_back(a); out.push_back(b);
        flatten(t->l, last + 1); flatten(t->r, last + 2);
    }
};
}  // namespace

void host_build_radiance_map(const float* surface_v, const float* surface_nrm, int n_surfaces, float area_per_sample,
                             std::vector<HostVolume>& volumes, std::vector<HostTreeElement>& tree) {
    volumes.clear(); tree.clear();
    // The reference draws volume positions with the process-wide rand(), never seeded (= srand(1)). random_r on a private
    // state reproduces exactly that glibc stream without touching (or racing on) the global generator, so every context
    // -- in any thread, built any number of times -- samples the same volumes, which multi-GPU replicas rely on.
    struct random_data rng; char rng_state[128];
    memset(&rng, 0, sizeof rng); memset(rng_state, 0, sizeof rng_state);
    initstate_r(1u, rng_state, sizeof rng_state, &rng);
    auto next_rand = [&]() { int32_t r = 0; random_r(&rng, &r); return (int)r; };
    for (int j = 0; j < n_surfaces; ++j) {
        const float* v = surface_v + 9 * j;
        int count = (int)std::floor(host_triangle_area(v) / area_per_sample);
        float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
        for (int i = 0; i < count; ++i) {
            HostVolume hv; float a1, a2;
            do {
                a1 = (float)next_rand() / (float)RAND_MAX; a2 = (float)next_rand() / (float)RAND_MAX;
                for (int k = 0; k < 3; ++k) hv.pos[k] = (v[k] + a1 * e1[k]) + a2 * e2[k];
            } while (a1 + a2 > 1.f);
            for (int k = 0; k < 3; ++k) hv.nrm[k] = surface_nrm[3 * j + k];
            hv.surface = j;
            volumes.push_back(hv);
        }
    }
    int nv = (int)volumes.size();
    std::vector<int> ids(nv); for (int i = 0; i < nv; ++i) ids[i] = i;
    Builder b{ volumes, tree, {} };
    Builder::Node* root = b.build(ids, 0);
    HostTreeElement r0{}; r0.dim = root->dim; r0.data = root->median;
    tree.push_back(r0);
    if (nv > 0) b.flatten(root, 0);
}


// ------------------------------------------------------------------------------------------------ candidate cells
// Why the lists are sufficient (DESIGN.md "Nearest volume"): let B be the cell's box (inflated by a rounding margin), v any
// volume of the class. Every point p of B has its closest same-class volume within R = min_v maxdist(v, B), so that
// volume has mindist(., B) <= R. Volumes farther than accept_r are never accepted by vcell_find, so only volumes with
// mindist <= min(R, accept_r) (plus slack for the float rounding of the device's distances, which also keeps every
// volume whose rounded distance could tie with the winner's) are listed.
namespace {
struct P3 { double x, y, z; };
static inline P3 sub(P3 a, P3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static inline double dotd(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// squared distance from p to triangle abc (Ericson, Real-Time Collision Detection 5.1.5)
static double point_triangle_dist2(P3 p, P3 a, P3 b, P3 c) {
    P3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
    double d1 = dotd(ab, ap), d2 = dotd(ac, ap);
    auto d2to = [&](P3 q) { P3 d = sub(p, q); return dotd(d, d); };
    if (d1 <= 0 && d2 <= 0) return d2to(a);
    P3 bp = sub(p, b); double d3 = dotd(ab, bp), d4 = dotd(ac, bp);
    if (d3 >= 0 && d4 <= d3) return d2to(b);
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); return d2to({ a.x + v * ab.x, a.y + v * ab.y, a.z + v * ab.z }); }
    P3 cp = sub(p, c); double d5 = dotd(ab, cp), d6 = dotd(ac, cp);
    if (d6 >= 0 && d5 <= d6) return d2to(c);
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); return d2to({ a.x + w * ac.x, a.y + w * ac.y, a.z + w * ac.z }); }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return d2to({ b.x + w * (c.x - b.x), b.y + w * (c.y - b.y), b.z + w * (c.z - b.z) });
    }
    double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
    return d2to({ a.x + ab.x * v + ac.x * w, a.y + ab.y * v + ac.y * w, a.z + ab.z * v + ac.z * w });
}
}  // namespace

void host_build_vcells(const float* sv, const int* sclass, int ns, const std::vector<HostVolume>& vol, const std::vector<int>& vclass,
                       const std::vector<HostTreeElement>& tree, float cell_h, float accept_r, float within_abs, HostVCells& out) {
    out = HostVCells{};
    const int nv = (int)vol.size();
    out.xtable.assign(4 * 2, -1); out.xcand.assign(12, 0.f);
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int i = 0; i < nv; ++i) for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], (double)vol[i].pos[k]); hi[k] = std::max(hi[k], (double)vol[i].pos[k]); }
    for (int i = 0; i < 9 * ns; ++i) { int k = i % 3; lo[k] = std::min(lo[k], (double)sv[i]); hi[k] = std::max(hi[k], (double)sv[i]); }
    const double ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    float h = std::max(cell_h, (float)(ext / 1000.0));                 // at most ~1000 cells per axis: the linear cell index stays below 2^31
    if (!(h > 0.f)) h = 1e-3f;
    out.h = h; out.ox = (float)lo[0] - h; out.oy = (float)lo[1] - h; out.oz = (float)lo[2] - h;
    const float inv_h = 1.f / h;
    out.nx = (int)grid_coord((float)hi[0], out.ox, inv_h) + 2; out.ny = (int)grid_coord((float)hi[1], out.oy, inv_h) + 2; out.nz = (int)grid_coord((float)hi[2], out.oz, inv_h) + 2;
    if (nv == 0 || !(accept_r > 0.f)) { out.table.assign(4 * 2, -1); out.cand.assign(16, 0.f); return; }
    const double scale = std::max(ext, 1.0), margin = 1e-5 * scale;      // box inflation: a point whose cell is decided by float rounding
    const double reach = (double)accept_r * (1.0 + 1e-4) + 1e-6 * scale;
    // coarse buckets of the volumes, cell size >= reach + box size, so a 3x3x3 block covers everything within reach of a box
    const double cs = std::max(reach + 2.0 * margin + (double)h, ext / 256.0);
    const int cnx = (int)std::floor((hi[0] - lo[0]) / cs) + 1, cny = (int)std::floor((hi[1] - lo[1]) / cs) + 1, cnz = (int)std::floor((hi[2] - lo[2]) / cs) + 1;
    auto ccell = [&](double x, int k, int n) { int c = (int)std::floor((x - lo[k]) / cs); return std::min(std::max(c, 0), n - 1); };
    std::vector<int> cstart((size_t)cnx * cny * cnz + 1, 0), corder(nv);
    {
        std::vector<int> cof(nv);
        for (int i = 0; i < nv; ++i) { cof[i] = (ccell(vol[i].pos[2], 2, cnz) * cny + ccell(vol[i].pos[1], 1, cny)) * cnx + ccell(vol[i].pos[0], 0, cnx); cstart[cof[i] + 1]++; }
        for (size_t k = 0; k + 1 < cstart.size(); ++k) cstart[k + 1] += cstart[k];
        std::vector<int> fill(cstart.begin(), cstart.end() - 1);
        for (int i = 0; i < nv; ++i) corder[fill[cof[i]]++] = i;
    }
    const bool vc_timing = getenv("RLPT_VC_TIMING") != nullptr; auto vc_t0 = std::chrono::steady_clock::now();
    auto vc_mark = [&](const char* what) { if (vc_timing) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "vcells %s: %.3f s\n", what, std::chrono::duration<double>(t - vc_t0).count()); vc_t0 = t; } };
    // (cell, class) pairs: every fine cell whose centre is within half a cell diagonal (+ margin) of a surface of that class
    std::unordered_map<uint64_t, int> keys;
    const double half_diag = 0.5 * std::sqrt(3.0) * (double)h + margin + 1e-4 * scale;
    for (int s = 0; s < ns; ++s) {
        const float* t = sv + 9 * (size_t)s;
        P3 a = { t[0], t[1], t[2] }, b = { t[3], t[4], t[5] }, c = { t[6], t[7], t[8] };
        int c0[3], c1[3];
        for (int k = 0; k < 3; ++k) {
            double mn = std::min(t[k], std::min(t[3 + k], t[6 + k])) - half_diag, mx = std::max(t[k], std::max(t[3 + k], t[6 + k])) + half_diag;
            const double o = k == 0 ? out.ox : (k == 1 ? out.oy : out.oz); const int n = k == 0 ? out.nx : (k == 1 ? out.ny : out.nz);
            c0[k] = std::max(0, (int)std::floor((mn - o) / h)); c1[k] = std::min(n - 1, (int)std::floor((mx - o) / h));
        }
        for (int z = c0[2]; z <= c1[2]; ++z) for (int y = c0[1]; y <= c1[1]; ++y) for (int x = c0[0]; x <= c1[0]; ++x) {
            P3 ctr = { out.ox + (x + 0.5) * (double)h, out.oy + (y + 0.5) * (double)h, out.oz + (z + 0.5) * (double)h };
            if (point_triangle_dist2(ctr, a, b, c) > half_diag * half_diag) continue;
            const uint64_t cell = (uint64_t)((size_t)(z * out.ny + y) * out.nx + x);
            keys.emplace((cell << 32) | (uint32_t)sclass[s], 0);
        }
    }
    vc_mark("keys");
    // candidate lists
    std::vector<uint64_t> order; order.reserve(keys.size());
    for (auto& kv : keys) order.push_back(kv.first);
    std::sort(order.begin(), order.end());
    struct Entry { int cell, cls, start, n4; };
    std::vector<Entry> entries; entries.reserve(order.size());
    std::vector<uint64_t> xkeys;
    // The lists of different (cell, class) pairs are independent: they are computed by all host threads (OpenMP) into per-pair vectors and
    // concatenated in key order afterwards, so the tables do not depend on the thread count. (Single-threaded this loop was 6.2 s of archway's
    // 6.5 s radiance-map build.)
    vc_mark("sort keys");
    const size_t nk = order.size();
    std::vector<std::vector<int>> lists(nk); std::vector<char> is_x(nk, 0);
#pragma omp parallel
    {
        std::vector<std::pair<double, int>> near;   // (mindist, volume)
#pragma omp for schedule(dynamic, 256)
        for (long long qi = 0; qi < (long long)nk; ++qi) {
            const uint64_t key = order[(size_t)qi];
            const int cell = (int)(key >> 32), cls = (int)(uint32_t)key;
            const int x = cell % out.nx, y = (cell / out.nx) % out.ny, z = cell / (out.nx * out.ny);
            const double blo[3] = { out.ox + x * (double)h - margin, out.oy + y * (double)h - margin, out.oz + z * (double)h - margin };
            const double bhi[3] = { blo[0] + h + 2 * margin, blo[1] + h + 2 * margin, blo[2] + h + 2 * margin };
            int k0[3], k1[3];
            k0[0] = ccell(blo[0] - reach, 0, cnx); k1[0] = ccell(bhi[0] + reach, 0, cnx);
            k0[1] = ccell(blo[1] - reach, 1, cny); k1[1] = ccell(bhi[1] + reach, 1, cny);
            k0[2] = ccell(blo[2] - reach, 2, cnz); k1[2] = ccell(bhi[2] + reach, 2, cnz);
            near.clear(); double R = 1e300;
            for (int cz = k0[2]; cz <= k1[2]; ++cz) for (int cy = k0[1]; cy <= k1[1]; ++cy) for (int cx = k0[0]; cx <= k1[0]; ++cx) {
                const size_t cc = (size_t)(cz * cny + cy) * cnx + cx;
                for (int j = cstart[cc]; j < cstart[cc + 1]; ++j) {
                    const int v = corder[j]; if (vclass[v] != cls) continue;
                    double mn2 = 0, mx2 = 0;
                    for (int k = 0; k < 3; ++k) {
                        const double q = vol[v].pos[k];
                        const double dlo = blo[k] - q, dhi = q - bhi[k];
                        const double dmin = std::max(0.0, std::max(dlo, dhi)), dmax = std::max(std::fabs(q - blo[k]), std::fabs(q - bhi[k]));
                        mn2 += dmin * dmin; mx2 += dmax * dmax;
                    }
                    const double mn = std::sqrt(mn2), mx = std::sqrt(mx2);
                    R = std::min(R, mx);
                    if (mn <= reach) near.push_back({ mn, v });
                }
            }
            if (near.empty() || R >= (double)accept_r * 0.999 - margin) is_x[(size_t)qi] = 1;     // some point of the cell may have no same-class volume within accept_r
            if (near.empty()) continue;
            const double lim = std::min(R * (1.0 + 1e-4) + 1e-6 * scale, reach);
            std::sort(near.begin(), near.end(), [](const std::pair<double, int>& p, const std::pair<double, int>& q) { return p.second < q.second; });
            std::vector<int>& l = lists[(size_t)qi];
            for (auto& pr : near) if (pr.first <= lim) l.push_back(pr.second);
        }
    }
    vc_mark("lists (parallel)");
    for (size_t qi = 0; qi < nk; ++qi) {
        const uint64_t key = order[qi];
        if (is_x[qi]) xkeys.push_back(key);
        const std::vector<int>& l = lists[qi];
        if (l.empty()) continue;
        Entry e{ (int)(key >> 32), (int)(uint32_t)key, (int)(out.cand.size() / 16), 0 };      // first group of 4 candidates
        int n = 0;
        for (int v : l) {
            float w; memcpy(&w, &v, 4);
            out.cand.push_back(vol[v].pos[0]); out.cand.push_back(vol[v].pos[1]); out.cand.push_back(vol[v].pos[2]); out.cand.push_back(w); ++n;
        }
        out.listed += (size_t)n;
        for (; n % 4 != 0; ++n) { const int v = -1; float w; memcpy(&w, &v, 4); out.cand.push_back(1e18f); out.cand.push_back(1e18f); out.cand.push_back(1e18f); out.cand.push_back(w); }
        e.n4 = n / 4; entries.push_back(e);
    }
    vc_mark("concatenate");
    if (out.cand.empty()) out.cand.assign(16, 0.f);
    out.keys = entries.size();
    // ---- second level. The kd search (kd_find) reaches leaf l from query p exactly when, for every ancestor whose split
    // separates them, |fl(p_k - split)| <= within_abs; per dimension the binding ancestor is the nearest bound of l's kd cell
    // [lo, hi] (lo = largest split with l in the right subtree, hi = smallest with l in the left), so
    //     visited(l, p)  <=>  for k = x, y, z:  fl(p_k - lo_k) >= -within_abs  and  fl(p_k - hi_k) <= within_abs.
    // A pair's list holds every same-class volume whose widened kd cell meets the cell's box, cut at the distance bound given
    // by a volume that is visited from EVERY point of the box.
    if (!xkeys.empty() && !tree.empty()) {
        std::vector<float> klo(3 * (size_t)nv, -INFINITY), khi(3 * (size_t)nv, INFINITY);
        {
            struct Item { unsigned idx; float lo[3], hi[3]; };
            std::vector<Item> stack; Item root{ 0u, { -INFINITY, -INFINITY, -INFINITY }, { INFINITY, INFINITY, INFINITY } }; stack.push_back(root);
            while (!stack.empty()) {
                Item it = stack.back(); stack.pop_back();
                const HostTreeElement& e = tree[it.idx];
                if (e.leaf) { const int v = (int)e.data; if (v >= 0 && v < nv) for (int k = 0; k < 3; ++k) { klo[3 * (size_t)v + k] = it.lo[k]; khi[3 * (size_t)v + k] = it.hi[k]; } continue; }
                Item l = it, r = it; l.idx = e.left; r.idx = e.right;
                l.hi[e.dim] = std::min(l.hi[e.dim], e.data); r.lo[e.dim] = std::max(r.lo[e.dim], e.data);
                stack.push_back(l); stack.push_back(r);
            }
        }
        int ncls = 0; for (int i = 0; i < nv; ++i) ncls = std::max(ncls, vclass[i] + 1);
        std::vector<std::vector<int>> by_class(std::max(ncls, 1));
        for (int i = 0; i < nv; ++i) if (vclass[i] >= 0) by_class[vclass[i]].push_back(i);
        const double w = (double)within_abs;
        struct XEntry { int cell, cls, start, n; };
        std::vector<XEntry> xent; xent.reserve(xkeys.size());
        out.xcand.clear();
        struct XC { double mn, mx; int v; bool always; };
        // independent per pair as well: all host threads, results concatenated in key order (2.8 of archway's remaining 3.9 s single-threaded)
        const size_t nx = xkeys.size();
        std::vector<std::vector<int>> xlists(nx); std::vector<char> xvalid(nx, 0);
#pragma omp parallel
        {
            std::vector<XC> poss;
#pragma omp for schedule(dynamic, 16)
            for (long long xi = 0; xi < (long long)nx; ++xi) {
                const uint64_t key = xkeys[(size_t)xi];
                const int cell = (int)(key >> 32), cls = (int)(uint32_t)key;
                if (cls < 0 || cls >= ncls) continue;
                xvalid[(size_t)xi] = 1;
                const int x = cell % out.nx, y = (cell / out.nx) % out.ny, z = cell / (out.nx * out.ny);
                const double blo[3] = { out.ox + x * (double)h - margin, out.oy + y * (double)h - margin, out.oz + z * (double)h - margin };
                const double bhi[3] = { blo[0] + h + 2 * margin, blo[1] + h + 2 * margin, blo[2] + h + 2 * margin };
                poss.clear(); double Rx = 1e300;
                for (int v : by_class[cls]) {
                    bool possible = true, always = true;
                    for (int k = 0; k < 3 && possible; ++k) {
                        const double lo = klo[3 * (size_t)v + k], hi = khi[3 * (size_t)v + k];
                        if (bhi[k] - lo < -w - margin || blo[k] - hi > w + margin) possible = false;
                        if (!(blo[k] - lo >= -w + margin && bhi[k] - hi <= w - margin)) always = false;
                    }
                    if (!possible) continue;
                    double mn2 = 0, mx2 = 0;
                    for (int k = 0; k < 3; ++k) {
                        const double q = vol[v].pos[k];
                        const double dmin = std::max(0.0, std::max(blo[k] - q, q - bhi[k])), dmax = std::max(std::fabs(q - blo[k]), std::fabs(q - bhi[k]));
                        mn2 += dmin * dmin; mx2 += dmax * dmax;
                    }
                    XC c{ std::sqrt(mn2), std::sqrt(mx2), v, always };
                    if (always) Rx = std::min(Rx, c.mx);
                    poss.push_back(c);
                }
                const double lim = Rx < 1e299 ? Rx * (1.0 + 1e-4) + 1e-6 * scale : 1e300;
                for (const XC& c : poss) if (c.mn <= lim) xlists[(size_t)xi].push_back(c.v);
            }
        }
        for (size_t xi = 0; xi < nx; ++xi) {
            if (!xvalid[xi]) continue;
            const uint64_t key = xkeys[xi];
            XEntry e{ (int)(key >> 32), (int)(uint32_t)key, (int)(out.xcand.size() / 12), 0 };
            for (int v : xlists[xi]) {
                float wv; memcpy(&wv, &v, 4);
                const float rec[12] = { vol[v].pos[0], vol[v].pos[1], vol[v].pos[2], wv, klo[3 * (size_t)v], klo[3 * (size_t)v + 1], klo[3 * (size_t)v + 2], khi[3 * (size_t)v],
                                        khi[3 * (size_t)v + 1], khi[3 * (size_t)v + 2], 0.f, 0.f };
                out.xcand.insert(out.xcand.end(), rec, rec + 12); ++e.n;
            }
            out.xlisted += (size_t)e.n;
            xent.push_back(e);                             // an empty list is an answer too: nothing of the class can be visited -> volume 0
        }
        if (out.xcand.empty()) out.xcand.assign(12, 0.f);
        out.xkeys = xent.size();
        size_t xs = 8; while (xs < 2 * xent.size() + 2) xs <<= 1;
        out.xtable.assign(4 * xs, -1);
        const uint32_t xmask = (uint32_t)(xs - 1);
        for (const XEntry& e : xent) {
            uint32_t hs = vcell_hash((uint32_t)e.cell, (uint32_t)e.cls) & xmask;
            while (out.xtable[4 * (size_t)hs] >= 0) hs = (hs + 1) & xmask;
            out.xtable[4 * (size_t)hs] = e.cell; out.xtable[4 * (size_t)hs + 1] = e.cls; out.xtable[4 * (size_t)hs + 2] = e.start; out.xtable[4 * (size_t)hs + 3] = e.n;
        }
    }
    vc_mark("second level");
    double fill = 3.0; if (const char* e = getenv("RLPT_VFILL")) fill = std::max(1.1, atof(e));
    size_t slots = 8; while ((double)slots < fill * (double)entries.size() + 2) slots <<= 1;
    out.table.assign(4 * slots, -1);
    const uint32_t mask = (uint32_t)(slots - 1);
    for (const Entry& e : entries) {
        uint32_t hs = vcell_hash((uint32_t)e.cell, (uint32_t)e.cls) & mask;
        while (out.table[4 * (size_t)hs] >= 0) hs = (hs + 1) & mask;
        out.table[4 * (size_t)hs] = e.cell; out.table[4 * (size_t)hs + 1] = e.cls; out.table[4 * (size_t)hs + 2] = e.start; out.table[4 * (size_t)hs + 3] = e.n4;
    }
    vc_mark("hash table");
}


void host_build_scan_units(const float* verts, int n_tri, HostScanUnits& out) {
    out = HostScanUnits{};
    auto vert = [&](int g) { return verts + 9 * (size_t)g; };
    double emax = 0.0, vmax = 0.0;
    for (int g = 0; g < n_tri; ++g) {
        const float* v = vert(g);
        for (int k = 0; k < 9; ++k) vmax = std::max(vmax, (double)std::fabs(v[k]));
        for (int k = 0; k < 3; ++k) { emax = std::max(emax, (double)std::fabs(v[3 + k] - v[k])); emax = std::max(emax, (double)std::fabs(v[6 + k] - v[k])); }
    }
    if (!std::isfinite(emax) || !std::isfinite(vmax)) return;
    const double tau = 1e-6 * std::max(1.0, vmax);          // how far a partner's vertices may be from the exact parallelogram
    std::vector<int> partner(n_tri, -1), apex(n_tri, 0); std::vector<char> used(n_tri, 0);
    for (int i = 0; i < n_tri; ++i) {
        if (used[i]) continue;
        const float* a = vert(i);
        for (int ap = 0; ap < 3 && partner[i] < 0; ++ap) {
            // the partner's vertices: the two of A other than the apex, and their sum minus the apex
            const int s1 = (ap + 1) % 3, s2 = (ap + 2) % 3;
            double want[3][3];
            for (int k = 0; k < 3; ++k) { want[0][k] = a[3 * s1 + k]; want[1][k] = a[3 * s2 + k]; want[2][k] = (double)a[3 * s1 + k] + (double)a[3 * s2 + k] - (double)a[3 * ap + k]; }
            for (int j = i + 1; j < n_tri && partner[i] < 0; ++j) {
                if (used[j]) continue;
                const float* b = vert(j);
                int hit = 0; bool taken[3] = { false, false, false };
                for (int w = 0; w < 3; ++w) for (int q = 0; q < 3; ++q) {
                    if (taken[q]) continue;
                    if (std::fabs(b[3 * q] - want[w][0]) <= tau && std::fabs(b[3 * q + 1] - want[w][1]) <= tau && std::fabs(b[3 * q + 2] - want[w][2]) <= tau) { taken[q] = true; ++hit; break; }
                }
                if (hit == 3) { partner[i] = j; apex[i] = ap; used[i] = used[j] = 1; }
            }
        }
    }
    for (int i = 0; i < n_tri; ++i) {
        if (partner[i] < 0) continue;
        const float* v = vert(i);
        const float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
        const double n[3] = { (double)e1[1] * e2[2] - (double)e1[2] * e2[1], (double)e1[2] * e2[0] - (double)e1[0] * e2[2], (double)e1[0] * e2[1] - (double)e1[1] * e2[0] };
        // where the partner lies in (u, v): u <= pu, v <= pv, u + v >= ps  (rlpt_device.cuh, unit_candidates)
        const float pu = apex[i] == 1 ? 0.f : 1.f, pv = apex[i] == 2 ? 0.f : 1.f, ps = apex[i] == 0 ? 1.f : 0.f;
        const float rec[16] = { v[0], v[1], v[2], e1[0], e1[1], e1[2], e2[0], e2[1], e2[2], (float)n[0], (float)n[1], (float)n[2], pu, pv, ps, 0.f };
        out.scan.insert(out.scan.end(), rec, rec + 16);
        out.slot_gid.push_back(i); out.slot_gid.push_back(partner[i]);
    }
    out.n_pairs = (int)(out.scan.size() / 16);
    // the remaining triangles: their own slots, and records without a partner (used by the bundle pre-test of the primary rays)
    for (int i = 0; i < n_tri; ++i) {
        if (used[i]) continue;
        out.slot_gid.push_back(i);
        const float* v = vert(i);
        const float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
        const double n[3] = { (double)e1[1] * e2[2] - (double)e1[2] * e2[1], (double)e1[2] * e2[0] - (double)e1[0] * e2[2], (double)e1[0] * e2[1] - (double)e1[1] * e2[0] };
        const float rec[16] = { v[0], v[1], v[2], e1[0], e1[1], e1[2], e2[0], e2[1], e2[2], (float)n[0], (float)n[1], (float)n[2], 0.f, 0.f, 0.f, 0.f };
        out.scan.insert(out.scan.end(), rec, rec + 16);
    }
    out.n_items = (int)(out.scan.size() / 16);
    // error-bound coefficients of unit_candidates: del = A (k1 Bm + k2), delx = k3 Bm
    out.k1 = (float)(2e-5 * emax); out.k2 = (float)(2e-5 * emax * emax + 8.0 * tau * emax); out.k3 = (float)(4e-5 * emax * emax); out.vmax = (float)vmax;
}

}  // namespace rlpt
