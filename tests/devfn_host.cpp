// tests/devfn_host.cpp -- host instantiation of the product's per-ray device functions (rlpt_device.cuh) and of the
// host-side radiance-map construction, for CPU tests of the host logic. TEST INFRASTRUCTURE: nothing in the product
// library runs these on the host. Built by tests/test_vcells_cpu.py with g++ -ffp-contract=off.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <map>
#include <array>
#include <vector>
#include <random>
#include "rlpt_radiance_host.h"
#include "rlpt_device.cuh"
using namespace rlpt;

static float within_abs_of(float max_dist) {
    const double md = (double)max_dist; float f = (float)std::sqrt(md);
    while (f > 0.f && (double)f * (double)f >= md) f = std::nextafterf(f, 0.f);
    for (float n = std::nextafterf(f, INFINITY); (double)n * (double)n < md; n = std::nextafterf(f, INFINITY)) f = n;
    return f;
}

// Builds the radiance map of the given surfaces, then answers `nq` seeded queries (points on the surfaces, jittered off
// them by `jitter`) twice: through the candidate cells (with the kd fallback, as find_volume does) and by the kd search
// alone. out[0] = mismatches, out[1] = queries decided by the cells, out[2] = volumes, out[3] = (cell, class) keys,
// out[4] = listed candidates, out[5] = table slots, out[6] = queries decided by the second level, out[7] / out[8] = its keys / candidates
extern "C" int devfn_check_vcells(const float* sv, int ns, float area_per_sample, float max_dist, float cell_factor, int nq, unsigned seed, float jitter, double* out) {
    std::vector<float> nrm(3 * (size_t)ns); std::vector<int> scls(ns);
    std::map<std::array<uint32_t, 3>, int> classes;
    for (int g = 0; g < ns; ++g) {
        host_triangle_normal(sv + 9 * (size_t)g, &nrm[3 * (size_t)g]);
        std::array<uint32_t, 3> key; for (int k = 0; k < 3; ++k) { float f = nrm[3 * g + k] == 0.f ? 0.f : nrm[3 * g + k]; memcpy(&key[k], &f, 4); }
        auto it = classes.find(key); if (it == classes.end()) it = classes.emplace(key, (int)classes.size()).first; scls[g] = it->second;
    }
    std::vector<HostVolume> vol; std::vector<HostTreeElement> tree;
    host_build_radiance_map(sv, nrm.data(), ns, area_per_sample, vol, tree);
    const int nv = (int)vol.size(), nt = (int)tree.size();
    std::vector<int> inner_of(nt, -1); int n_inner = 0;
    for (int i = 0; i < nt; ++i) if (!tree[i].leaf) inner_of[i] = n_inner++;
    auto child_word = [&](unsigned idx) -> uint32_t { return tree[idx].leaf ? (KD_LEAF | (uint32_t)(int)tree[idx].data) : (uint32_t)inner_of[idx]; };
    struct Inner { float split; uint32_t l, r; int dim; };
    std::vector<Inner> kd(std::max(n_inner, 1));
    for (int i = 0; i < nt; ++i) if (!tree[i].leaf) kd[inner_of[i]] = Inner{ tree[i].data, child_word(tree[i].left), child_word(tree[i].right), tree[i].dim };
    std::vector<int> vcls(nv); for (int i = 0; i < nv; ++i) vcls[i] = scls[vol[i].surface];
    const float within = within_abs_of(max_dist), accept = within * (1.f - 1e-5f);
    HostVCells hv; host_build_vcells(sv, scls.data(), ns, vol, vcls, tree, cell_factor * std::sqrt(area_per_sample), accept, within, hv);
    const uint32_t xmask = (uint32_t)(hv.xtable.size() / 4 - 1);
    VCells g{}; g.ox = hv.ox; g.oy = hv.oy; g.oz = hv.oz; g.inv_h = 1.f / hv.h; g.nx = hv.nx; g.ny = hv.ny; g.nz = hv.nz; g.mask = (uint32_t)(hv.table.size() / 4 - 1); g.accept_r = accept;
    const uint32_t root = child_word(0); const float rx = tree[0].pos[0], ry = tree[0].pos[1], rz = tree[0].pos[2];
    auto kd_only = [&](float px, float py, float pz, int cls) {
        return kd_find([&](uint32_t idx, float& split, uint32_t& l, uint32_t& r, int& dim) { split = kd[idx].split; l = kd[idx].l; r = kd[idx].r; dim = kd[idx].dim; },
                       [&](int v, float& x, float& y, float& z, int& c) { x = vol[v].pos[0]; y = vol[v].pos[1]; z = vol[v].pos[2]; c = vcls[v]; },
                       root, rx, ry, rz, px, py, pz, cls, within);
    };
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(0.f, 1.f); std::normal_distribution<float> Nrm(0.f, 1.f);
    double mism = 0, decided = 0, decided2 = 0;
    for (int q = 0; q < nq; ++q) {
        const int s = (int)(rng() % (unsigned)ns); const float* t = sv + 9 * (size_t)s;
        float u = U(rng), v = U(rng); if (u + v > 1.f) { u = 1.f - u; v = 1.f - v; }
        float p[3]; for (int k = 0; k < 3; ++k) p[k] = t[k] + u * (t[3 + k] - t[k]) + v * (t[6 + k] - t[k]) + jitter * Nrm(rng);
        const int cls = scls[s];
        const float d0 = kd_distance(p[0], p[1], p[2], rx, ry, rz);
        int a = vcell_find(g, [&](uint32_t i, int& c, int& k, int& st, int& n) { c = hv.table[4 * (size_t)i]; k = hv.table[4 * (size_t)i + 1]; st = hv.table[4 * (size_t)i + 2]; n = hv.table[4 * (size_t)i + 3]; },
                           [&](int i, float& x, float& y, float& z, int& vv) { x = hv.cand[4 * (size_t)i]; y = hv.cand[4 * (size_t)i + 1]; z = hv.cand[4 * (size_t)i + 2]; memcpy(&vv, &hv.cand[4 * (size_t)i + 3], 4); },
                           p[0], p[1], p[2], cls, d0);
        if (a < 0) {
            a = vext_find(g, xmask, [&](uint32_t i, int& c, int& k, int& st, int& n) { c = hv.xtable[4 * (size_t)i]; k = hv.xtable[4 * (size_t)i + 1]; st = hv.xtable[4 * (size_t)i + 2]; n = hv.xtable[4 * (size_t)i + 3]; },
                          [&](int i, float& x, float& y, float& z, int& vv, float (&lo)[3], float (&hi)[3]) {
                              const float* r = &hv.xcand[12 * (size_t)i]; x = r[0]; y = r[1]; z = r[2]; memcpy(&vv, &r[3], 4);
                              lo[0] = r[4]; lo[1] = r[5]; lo[2] = r[6]; hi[0] = r[7]; hi[1] = r[8]; hi[2] = r[9];
                          }, p[0], p[1], p[2], cls, d0, within);
            if (a >= 0) decided2 += 1;
        }
        const int b = kd_only(p[0], p[1], p[2], cls);
        if (a >= 0) { decided += 1; if (a != b) mism += 1; }
    }
    out[0] = mism; out[1] = decided; out[2] = nv; out[3] = (double)hv.keys; out[4] = (double)hv.listed; out[5] = (double)(hv.table.size() / 4);
    out[6] = decided2; out[7] = (double)hv.xkeys; out[8] = (double)hv.xlisted;
    return 0;
}

// Conservativeness of the brute-force pre-test (unit_candidates): `nr` seeded rays (origins on the surfaces or at the camera,
// directions uniform on the sphere); every primitive the exact solve (tri_solve, the reference's arithmetic) accepts must be
// marked. out[0] = accepted-but-unmarked (must be 0), out[1] = marked per ray, out[2] = accepted per ray, out[3] = units,
// out[4] = parallelogram pairs
extern "C" int devfn_check_units(const float* verts, int n_tri, int nr, unsigned seed, float H, const float* cam, double* out) {
    HostScanUnits hu; host_build_scan_units(verts, n_tri, hu);
    const int nu = hu.n_pairs;
    if ((int)hu.slot_gid.size() != n_tri) return 1;
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(0.f, 1.f); std::normal_distribution<float> Nrm(0.f, 1.f);
    double missed = 0, marked = 0, accepted = 0, pairs = nu;
    for (int q = 0; q < nr; ++q) {
        float o[3], d[3];
        if (q % 4 == 0) { o[0] = cam[0]; o[1] = cam[1]; o[2] = cam[2]; }
        else {
            const float* t = verts + 9 * (size_t)(rng() % (unsigned)n_tri);
            float u = U(rng), v = U(rng); if (u + v > 1.f) { u = 1.f - u; v = 1.f - v; }
            for (int k = 0; k < 3; ++k) o[k] = t[k] + u * (t[3 + k] - t[k]) + v * (t[6 + k] - t[k]);
        }
        for (int k = 0; k < 3; ++k) d[k] = Nrm(rng);
        if (q % 7 == 0) d[rng() % 3] = 0.f;                  // axis-aligned planes: grazing and edge-on rays
        f3 dn = normalize_ref(f3{ d[0], d[1], d[2] });
        if (q % 4 != 0) for (int k = 0; k < 3; ++k) o[k] = fmaf(RAY_EPS, (&dn.x)[k], o[k]);
        const float sdx = dn.x * H, sdy = dn.y * H, sdz = dn.z * H, a0 = 0.f - sdx, a1 = 0.f - sdy, a2 = 0.f - sdz;
        const float A_ = fmaxf(fabsf(a0), fmaxf(fabsf(a1), fabsf(a2))), Bm = fmaxf(fabsf(o[0]), fmaxf(fabsf(o[1]), fabsf(o[2]))) + hu.vmax;
        const float del = A_ * (hu.k1 * Bm + hu.k2), kx = (3.f * del) / (hu.k3 * Bm);
        std::vector<char> mark(n_tri, 0);
        for (int u = 0; u < nu; ++u) {
            const float* r = &hu.scan[16 * (size_t)u];
            UnitRec rec{ r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14] };
            unsigned bits = unit_candidates(rec, o[0], o[1], o[2], a0, a1, a2, del, kx);
            if (bits & 1u) mark[hu.slot_gid[2 * u]] = 1;
            if (bits & 2u) mark[hu.slot_gid[2 * u + 1]] = 1;
        }
        for (int sl = 2 * nu; sl < n_tri; ++sl) {
            const int g = hu.slot_gid[sl]; const float* v = verts + 9 * (size_t)g;
            const float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
            TriRec tr{ v[0], v[1], v[2], e1[0], e1[1], e1[2], e2[0], e2[1], e2[2], fmaf(e1[1], e2[2], -(e1[2] * e2[1])) };
            if (tri_candidate_small(tr, o[0], o[1], o[2], a0, a1, a2)) mark[g] = 1;
        }
        for (int g = 0; g < n_tri; ++g) {
            const float* v = verts + 9 * (size_t)g;
            const float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
            TriRec tr{ v[0], v[1], v[2], e1[0], e1[1], e1[2], e2[0], e2[1], e2[2], fmaf(e1[1], e2[2], -(e1[2] * e2[1])) };
            float t;
            const bool acc = tri_solve(tr, o[0], o[1], o[2], a0, a1, a2, T_MISS, t);
            if (mark[g]) marked += 1;
            if (acc) { accepted += 1; if (!mark[g]) missed += 1; }
        }
    }
    out[0] = missed; out[1] = marked / nr; out[2] = accepted / nr; out[3] = nu; out[4] = pairs;
    return 0;
}
