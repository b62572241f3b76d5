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
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>

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

}  // namespace rlpt
