// tests/exact_forms_host.cpp -- TEST INFRASTRUCTURE. Three places where the product evaluates something in a cheaper form than
// the straightforward one, each checked here against the straightforward form written out, on the host instantiation of the
// product's own functions (rlpt_device.cuh): results must be identical bit for bit.
//   1. square_to_hemisphere: selects + one division vs. the four-way branch
//   2. tri_solve: early outs and the "surely inside" shortcut vs. the reference's three divisions (G/rays/ray.cu:39-74,115-141)
//   3. the zero-contribution statistic: sum <= threshold vs. mean(rgb) < THROUGHPUT_THRESHOLD with its division by three
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include "rlpt_device.cuh"
using namespace rlpt;

static void branchy_s2h(float sx, float sy, float& xh, float& yh, float& zh) {
    float a = 2.f * sx - 1.f, b = 2.f * sy - 1.f, r, phi;
    const float q = 0.78539816339744830962f;
    if (a > -b) { if (a > b) { r = a; phi = q * (b / a); } else { r = b; phi = q * (2.f - a / b); } }
    else { if (a < b) { r = -a; phi = q * (4.f + b / a); } else { r = -b; phi = (b != 0.f) ? q * (6.f - a / b) : 0.f; } }
    float s = sinf(phi), c = cosf(phi), st = r * sqrtf(2.f - r * r);
    xh = st * c; yh = 1.f - r * r; zh = st * s;
}
extern "C" long exact_check_hemisphere(int n_random, int grid, unsigned seed) {
    std::mt19937 g(seed); std::uniform_real_distribution<float> U(0.f, 1.f);
    long bad = 0;
    auto check = [&](float x, float y) { float a[3], b[3]; branchy_s2h(x, y, a[0], a[1], a[2]); square_to_hemisphere(x, y, b[0], b[1], b[2]); bad += memcmp(a, b, 12) != 0; };
    for (int i = 0; i < n_random; ++i) check(U(g), U(g));
    for (int i = 0; i <= grid; ++i) for (int j = 0; j <= grid; ++j) check(i / (float)grid, j / (float)grid);        // cell corners, diagonals, axes, the centre
    return bad;
}

static bool plain_solve(const TriRec& r, float ox, float oy, float oz, float a0, float a1, float a2, float& t) {
    float T2 = RLPT_FMA(r.e2z, a1, -RLPT_MUL(r.e2y, a2)), T3 = RLPT_FMA(r.e1z, a1, -RLPT_MUL(r.e1y, a2));
    float detA = RLPT_FMA(r.e2x, T3, RLPT_FMA(a0, r.T1, -RLPT_MUL(r.e1x, T2)));
    float bx = RLPT_SUB(ox, r.v0x), by = RLPT_SUB(oy, r.v0y), bz = RLPT_SUB(oz, r.v0z);
    float p75 = RLPT_MUL(r.e1z, by), p80 = RLPT_MUL(r.e1y, bz);
    float U2 = RLPT_FMA(r.e2z, by, -RLPT_MUL(r.e2y, bz)), V3 = RLPT_FMA(a1, bz, -RLPT_MUL(a2, by));
    float dy = RLPT_FMA(r.e2x, V3, RLPT_FMA(a0, U2, -RLPT_MUL(T2, bx)));
    float dz = RLPT_FMA(T3, bx, RLPT_FMA(a0, RLPT_SUB(p80, p75), -RLPT_MUL(r.e1x, V3)));
    if (!(detA != 0.f)) return false;
    float dx = RLPT_FMA(r.e2x, RLPT_SUB(p75, p80), RLPT_FMA(r.T1, bx, -RLPT_MUL(r.e1x, U2)));
    float u = RLPT_DIV(dy, detA), v = RLPT_DIV(dz, detA);
    if (!(u >= 0.f && v >= 0.f && RLPT_ADD(u, v) <= 1.f)) return false;
    t = RLPT_DIV(dx, detA);
    return t >= 0.f;
}
// rays aimed at the inside, the three edges, the vertices and past the triangle, from random origins; out[0] = solves, out[1] =
// accepted by the plain form, returns the number of disagreements (a hit the product drops because it is farther than best_t is
// not one: the caller could not have kept it)
extern "C" long exact_check_tri_solve(int n_tri, int n_ray, unsigned seed, double* out) {
    std::mt19937 g(seed); std::uniform_real_distribution<float> U(-1.f, 1.f), W(0.f, 1.f);
    long n = 0, bad = 0, acc = 0;
    for (int tri = 0; tri < n_tri; ++tri) {
        float v0[3], v1[3], v2[3];
        const float sc = tri % 3 == 0 ? 0.02f : (tri % 3 == 1 ? 0.3f : 2.f);
        for (int k = 0; k < 3; ++k) { v0[k] = U(g); v1[k] = v0[k] + sc * U(g); v2[k] = v0[k] + sc * U(g); }
        if (tri % 5 == 0) { v1[1] = v0[1]; v2[1] = v0[1]; }                            // axis-aligned planes like the Cornell walls
        TriRec r{ v0[0], v0[1], v0[2], v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2], v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2], 0.f };
        r.T1 = RLPT_FMA(r.e1y, r.e2z, -RLPT_MUL(r.e1z, r.e2y));
        for (int ray = 0; ray < n_ray; ++ray) {
            float o[3] = { U(g), U(g), U(g) }, p[3], bu, bv; const int mode = ray % 6;
            if (mode == 0) { bu = W(g); bv = W(g) * (1 - bu); }
            else if (mode == 1) { bu = W(g); bv = 1 - bu; }
            else if (mode == 2) { bu = 0; bv = W(g); }
            else if (mode == 3) { bu = W(g); bv = 0; }
            else if (mode == 4) { bu = (float)((ray / 6) % 2); bv = (float)((ray / 12) % 2) * (1 - bu); }
            else { bu = 2 * U(g); bv = 2 * U(g); }
            for (int k = 0; k < 3; ++k) p[k] = v0[k] + bu * (v1[k] - v0[k]) + bv * (v2[k] - v0[k]);
            if (ray % 7 == 0) for (int k = 0; k < 3; ++k) p[k] += 1e-7f * U(g);
            f3 dn = normalize_ref(f3{ p[0] - o[0], p[1] - o[1], p[2] - o[2] });
            const float H = 512.f, a0 = RLPT_SUB(0.f, RLPT_MUL(dn.x, H)), a1 = RLPT_SUB(0.f, RLPT_MUL(dn.y, H)), a2 = RLPT_SUB(0.f, RLPT_MUL(dn.z, H));
            for (int bt = 0; bt < 2; ++bt) {
                const float best = bt ? 0.002f : T_MISS;
                float t0 = -1, t1 = -1;
                const bool h0 = plain_solve(r, o[0], o[1], o[2], a0, a1, a2, t0), h1 = tri_solve(r, o[0], o[1], o[2], a0, a1, a2, best, t1);
                const bool same = (h0 == h1 && (!h0 || memcmp(&t0, &t1, 4) == 0)) || (h0 && !h1 && t0 > best);
                ++n; acc += h0; bad += !same;
            }
        }
    }
    out[0] = (double)n; out[1] = (double)acc;
    return bad;
}

// every float in [lo, hi] (walked with nextafter) plus n_random sums of three: (x / 3 < 0.0001f) == (x <= threshold)
extern "C" long exact_check_zero_contribution(float threshold, float lo, float hi, int n_random, unsigned seed) {
    long bad = 0;
    for (float x = lo; x <= hi; x = nextafterf(x, 1.f)) bad += ((x / 3.f < 0.0001f) != (x <= threshold));
    std::mt19937 g(seed); std::uniform_real_distribution<float> U(0.f, 4e-4f);
    for (int i = 0; i < n_random; ++i) { float a = U(g), b = U(g) * 0.3f, c = U(g) * 0.1f, x = a + b + c; bad += ((x / 3.f < 0.0001f) != (x <= threshold)); }
    const float special[] = { 0.f, -0.f, 1.f, INFINITY, NAN, 1e-30f };
    for (float x : special) bad += ((x / 3.f < 0.0001f) != (x <= threshold));
    return bad;
}
