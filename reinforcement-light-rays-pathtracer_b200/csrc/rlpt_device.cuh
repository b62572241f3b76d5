// rlpt_device.cuh -- per-ray device functions of the hot path (sm_100a). Everything here is FP32 scalar math on the
// FP32 pipes; no tensor-core work belongs here (BASELINE.json north_star).
//
// Bit-exactness: the closest-hit test, the ray normalisation and the nearest-volume distance reproduce the exact
// rounding sequence of the reference's kernels as nvcc 12.9 builds them for sm_100a (-fmad=true); the sequences were
// read from the SASS of G/rays/ray.cu and G/radiance_volumes/radiance_map.cu and are spelled out with explicit
// round-to-nearest intrinsics so no compiler setting can re-associate or re-contract them ("Bit-exact
// arithmetic"). The functions are __host__ __device__ only so that tests/ can also exercise them on the CPU through
// tests/devfn_host.cpp; the product library never runs them on the host.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RLPT_HD __host__ __device__ __forceinline__
#else
#define RLPT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define RLPT_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define RLPT_MUL(a, b) __fmul_rn((a), (b))
#define RLPT_ADD(a, b) __fadd_rn((a), (b))
#define RLPT_SUB(a, b) __fsub_rn((a), (b))
#define RLPT_DIV(a, b) __fdiv_rn((a), (b))
#define RLPT_SQRT(a) __fsqrt_rn((a))
#else
// host instantiation (tests only): compiled with -ffp-contract=off so each operator is one rounding
#define RLPT_FMA(a, b, c) fmaf((a), (b), (c))
#define RLPT_MUL(a, b) ((a) * (b))
#define RLPT_ADD(a, b) ((a) + (b))
#define RLPT_SUB(a, b) ((a) - (b))
#define RLPT_DIV(a, b) ((a) / (b))
#define RLPT_SQRT(a) sqrtf((a))
#endif

namespace rlpt {

constexpr int GRID = 12;
constexpr int CELLS = 144;
constexpr float RHO = 1.f / (2.f * 3.1415926535f);        // G/constants/image_settings.h:13
constexpr float GRID_RHO = 1.f / 144.f;                   // G/constants/radiance_volumes_settings.h:10
constexpr float PI_F = 3.14159265358979323846f;           // (float)M_PI
constexpr float T_MISS = 999999.f;                        // G/rays/ray.cu:18
constexpr float RAY_EPS = 0.00001f;                       // G/path_tracing/default_path_tracing.cu:80

struct f3 { float x, y, z; };

// ---------------------------------------------------------------- RNG: Philox4x32-10, counter = (pixel, sample, bounce, purpose)
// Replaces the per-pixel XORWOW state array of init_rand_state (G/utils/cuda_helpers.cu:16-25): no state in HBM, and
// a path's random numbers do not depend on which GPU or which queue slot traces it.
enum { PURPOSE_CAMERA = 0, PURPOSE_BOUNCE = 1 };
RLPT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
RLPT_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// cuRAND's curand_uniform convention, (0,1]
RLPT_HD float u01(uint32_t x) { return x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }
RLPT_HD void draw4(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t purpose, float& u0, float& u1, float& u2, float& u3) {
    uint32_t c0 = pixel, c1 = sample, c2 = bounce, c3 = purpose;
    philox4x32_10(seed, 0u, c0, c1, c2, c3);
    u0 = u01(c0); u1 = u01(c1); u2 = u01(c2); u3 = u01(c3);
}

// ---------------------------------------------------------------- ray setup
// glm::normalize as Ray::Ray applies it (G/rays/ray.cu:6-14): v * (1/sqrt(dot)), dot = fma(z,z,fma(x,x,y*y)).
RLPT_HD f3 normalize_ref(f3 v) {
    float d = RLPT_FMA(v.z, v.z, RLPT_FMA(v.x, v.x, RLPT_MUL(v.y, v.y)));
    float inv = RLPT_DIV(1.f, RLPT_SQRT(d));
    return { RLPT_MUL(v.x, inv), RLPT_MUL(v.y, inv), RLPT_MUL(v.z, inv) };
}

// Ray::sample_ray_through_pixel + rotate_ray (G/rays/ray.cu:144-172). cy/sy/cx/sx = cos/sin of yaw_y, yaw_x.
RLPT_HD f3 camera_dir(int px, int py, float u0, float u1, int width, int height, bool rotated, float cy, float sy, float cx, float sx) {
    float x = (float)px + u0, y = (float)py + u1;
    f3 d = normalize_ref(f3{ x - (float)width / 2.f, y - (float)height / 2.f, (float)height });
    if (rotated) {
        f3 r1 = { (cy * d.x + 0.f * d.y) + (-sy) * d.z, d.y, (sy * d.x + 0.f * d.y) + cy * d.z };
        d = f3{ r1.x, (cx * r1.y) + sx * r1.z, ((-sx) * r1.y) + cx * r1.z };
    }
    return d;
}

// ---------------------------------------------------------------- ray / triangle (G/rays/ray.cu:39-74,115-141)
// Triangle record, three float4: q0 = (v0.x, v0.y, v0.z, e1.x), q1 = (e1.y, e1.z, e2.x, e2.y), q2 = (e2.z, T1, -, -)
// with e1 = v1-v0, e2 = v2-v0 and T1 = fma(e1.y, e2.z, -(e1.z*e2.y)) -- all ray-independent single roundings of the
// reference's own expressions, so storing them changes no bit of the result.
struct TriRec { float v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z, T1; };

// One Cramer solve. a = -(dir*SCREEN_HEIGHT) (the first column of A), o = ray origin.
// Returns true when the reference's acceptance test passes; t is the reference's solution.x.
RLPT_HD bool tri_solve(const TriRec& r, float ox, float oy, float oz, float a0, float a1, float a2, float best_t, float& t) {
    float T2 = RLPT_FMA(r.e2z, a1, -RLPT_MUL(r.e2y, a2));
    float T3 = RLPT_FMA(r.e1z, a1, -RLPT_MUL(r.e1y, a2));
    float detA = RLPT_FMA(r.e2x, T3, RLPT_FMA(a0, r.T1, -RLPT_MUL(r.e1x, T2)));
    float bx = RLPT_SUB(ox, r.v0x), by = RLPT_SUB(oy, r.v0y), bz = RLPT_SUB(oz, r.v0z);
    float p75 = RLPT_MUL(r.e1z, by), p80 = RLPT_MUL(r.e1y, bz);
    float U2 = RLPT_FMA(r.e2z, by, -RLPT_MUL(r.e2y, bz));
    float V3 = RLPT_FMA(a1, bz, -RLPT_MUL(a2, by));
    float dy = RLPT_FMA(r.e2x, V3, RLPT_FMA(a0, U2, -RLPT_MUL(T2, bx)));
    float dz = RLPT_FMA(T3, bx, RLPT_FMA(a0, RLPT_SUB(p80, p75), -RLPT_MUL(r.e1x, V3)));
    if (!(detA != 0.f)) return false;
    // Exact-safe early outs: each rejects only what the reference's own rounded quotients reject too; everything else
    // goes through the reference's divisions below.
    // (1) rn(n/d) is negative exactly when n and d have strictly opposite signs and the quotient does not round to -0,
    //     which needs |n/d| > 2^-150; for |d| < 2^23 that holds whenever n*d < -2^-100.
    const float kTiny = -7.888609052210118e-31f;   // -2^-100
    const float ad = fabsf(detA);
    if (ad < 8388608.f) {
        if (RLPT_MUL(dy, detA) < kTiny || RLPT_MUL(dz, detA) < kTiny) return false;
    }
    // (2) u + v > 1: with a = dy/detA, b = dz/detA, |dy + dz| > |detA| (1 + 1e-6) in floats means a + b > 1 + 8e-7 when
    //     a, b >= 0 (and b > 1 + 8e-7 when the other quotient rounds to -0), so rn(rn(a) + rn(b)) > 1: three roundings
    //     of relative error 2^-24 cannot bring it back to 1. Opposite-sign pairs never trigger it wrongly: they make the
    //     sum smaller, not larger. NaN/inf compare false and fall through.
    if (fabsf(RLPT_ADD(dy, dz)) > RLPT_MUL(ad, 1.000001f)) return false;
    // (3) farther than the current best: |dx| > best_t |detA| (1 + 1e-6) means rn(dx/detA) > best_t or < 0; the caller
    //     keeps a hit only when t < best_t, or t == best_t with a lower primitive id
    float dx = RLPT_FMA(r.e2x, RLPT_SUB(p75, p80), RLPT_FMA(r.T1, bx, -RLPT_MUL(r.e1x, U2)));
    //     (not used while best_t == 0: a quotient that underflows to -0 ties with it)
    const float lim = RLPT_MUL(RLPT_MUL(best_t, ad), 1.000001f);
    if (lim > 0.f && fabsf(dx) > lim) return false;
    // (4) surely inside: with s = sign(detA), s dy >= 0 and s dz >= 0 make both quotients >= 0 (or -0, which passes too), and
    //     s (dy + dz) <= |detA| (1 - 1e-6) in floats means a + b <= 1 - 8e-7 exactly, so rn(rn(a) + rn(b)) <= 1 after its three
    //     roundings. Then u and v need not be divided out at all -- two of the three divisions, 4 % of k_isect's instructions.
    //     The 1e-6 band around the edges, NaN/inf and |detA| >= 2^23 go through the reference's own expressions.
    const float sgn = detA < 0.f ? -1.f : 1.f;
    const float ys = RLPT_MUL(dy, sgn), zs = RLPT_MUL(dz, sgn);
    if (!(ad < 8388608.f && ys >= 0.f && zs >= 0.f && RLPT_ADD(ys, zs) <= RLPT_MUL(ad, 0.999999f))) {
        float u = RLPT_DIV(dy, detA), v = RLPT_DIV(dz, detA);
        if (!(u >= 0.f && v >= 0.f && RLPT_ADD(u, v) <= 1.f)) return false;
    }
    t = RLPT_DIV(dx, detA);
    return t >= 0.f;
}

// Phase 1 of the two-phase brute-force scan: the determinant and the two barycentric numerators only, and early outs
// (1) and (2) of tri_solve on them -- branch-free, no division. Returns true when tri_solve has to be run for this
// triangle; it returns false only where tri_solve would return false before reaching its divisions.
RLPT_HD bool tri_candidate(const TriRec& r, float ox, float oy, float oz, float a0, float a1, float a2) {
    float T2 = RLPT_FMA(r.e2z, a1, -RLPT_MUL(r.e2y, a2));
    float T3 = RLPT_FMA(r.e1z, a1, -RLPT_MUL(r.e1y, a2));
    float detA = RLPT_FMA(r.e2x, T3, RLPT_FMA(a0, r.T1, -RLPT_MUL(r.e1x, T2)));
    float bx = RLPT_SUB(ox, r.v0x), by = RLPT_SUB(oy, r.v0y), bz = RLPT_SUB(oz, r.v0z);
    float p75 = RLPT_MUL(r.e1z, by), p80 = RLPT_MUL(r.e1y, bz);
    float U2 = RLPT_FMA(r.e2z, by, -RLPT_MUL(r.e2y, bz));
    float V3 = RLPT_FMA(a1, bz, -RLPT_MUL(a2, by));
    float dy = RLPT_FMA(r.e2x, V3, RLPT_FMA(a0, U2, -RLPT_MUL(T2, bx)));
    float dz = RLPT_FMA(T3, bx, RLPT_FMA(a0, RLPT_SUB(p80, p75), -RLPT_MUL(r.e1x, V3)));
    const float kTiny = -7.888609052210118e-31f;   // -2^-100
    const float ad = fabsf(detA);
    const bool sign_out = ad < 8388608.f && fminf(RLPT_MUL(dy, detA), RLPT_MUL(dz, detA)) < kTiny;
    const bool sum_out = fabsf(RLPT_ADD(dy, dz)) > RLPT_MUL(ad, 1.000001f);
    return detA != 0.f && !sign_out && !sum_out;
}

// tri_candidate without its two guards, for scenes where they cannot fire (checked on the host, SceneDev::det_small):
//   * |detA| < 2^23 always: detA = a . (e1 x e2) with |a| = SCREEN_HEIGHT |dir|, so |detA| <= H |e1| |e2| (1 + 1e-6)
//   * detA == 0: then dy detA = dz detA = 0, no sign early-out fires, and the triangle is at worst a false candidate that
//     tri_solve rejects -- conservative, never wrong
RLPT_HD bool tri_candidate_small(const TriRec& r, float ox, float oy, float oz, float a0, float a1, float a2) {
    float T2 = RLPT_FMA(r.e2z, a1, -RLPT_MUL(r.e2y, a2));
    float T3 = RLPT_FMA(r.e1z, a1, -RLPT_MUL(r.e1y, a2));
    float detA = RLPT_FMA(r.e2x, T3, RLPT_FMA(a0, r.T1, -RLPT_MUL(r.e1x, T2)));
    float bx = RLPT_SUB(ox, r.v0x), by = RLPT_SUB(oy, r.v0y), bz = RLPT_SUB(oz, r.v0z);
    float p75 = RLPT_MUL(r.e1z, by), p80 = RLPT_MUL(r.e1y, bz);
    float U2 = RLPT_FMA(r.e2z, by, -RLPT_MUL(r.e2y, bz));
    float V3 = RLPT_FMA(a1, bz, -RLPT_MUL(a2, by));
    float dy = RLPT_FMA(r.e2x, V3, RLPT_FMA(a0, U2, -RLPT_MUL(T2, bx)));
    float dz = RLPT_FMA(T3, bx, RLPT_FMA(a0, RLPT_SUB(p80, p75), -RLPT_MUL(r.e1x, V3)));
    const float kTiny = -7.888609052210118e-31f;   // -2^-100
    const bool sign_out = fminf(RLPT_MUL(dy, detA), RLPT_MUL(dz, detA)) < kTiny;
    const bool sum_out = fabsf(RLPT_ADD(dy, dz)) > RLPT_MUL(fabsf(detA), 1.000001f);
    return !(sign_out || sum_out);
}

// Conservative pre-test of one scan unit: a triangle A = (v0; e1, e2) and optionally the triangle B that completes the
// parallelogram v0 + u e1 + v e2, u, v in [0, 1] (B = {v0 + e1, v0 + e2, v0 + e1 + e2}). With a = -(dir * SCREEN_HEIGHT),
// b = o - v0, n = e1 x e2 and c = a x b, Cramer's numerators of the reference's system (G/rays/ray.cu:39-74,115-141) are
//     D = det[a e1 e2] = a . n      X = det[b e1 e2] = b . n (t)      Y = det[a b e2] = e2 . c (u)      Z = det[a e1 b] = -(e1 . c) (v)
// -- 21 multiply-adds for BOTH triangles instead of the reference's expression tree per triangle. They are NOT the
// reference's roundings, so they decide nothing: they only discard triangles that the exact solve (tri_solve) is certain to
// reject. Error bounds (u = 2^-24; A = max |a_i|, E = largest |edge component| of the scene, Bm = max |o_i| + largest
// |vertex coordinate|): the reference's own dy, dz are within 36 u A E Bm of the exact values, these Y, Z too; detA and D
// within 30 u A E^2; dx and X within 36 u E^2 Bm. del = 2e-5 A E (Bm + E) and delx = 4e-5 E^2 Bm are more than twice
// the sums (plus the slack of a parallelogram whose fourth vertex is off by up to 1e-6 of the scene size). A triangle is
// kept unless, by more than these bounds, t < 0, or a barycentric numerator has the wrong sign, or u + v is on the wrong
// side of 1 (all with the one threshold 3 del); when |D| <= del the orientation itself is uncertain and everything is kept.
// Returns bit 0 = A, bit 1 = B. Triangles without a parallelogram partner go through tri_candidate(_small) instead.
// The partner B shares an edge of A and completes a parallelogram; which edge decides where B lies in A's (u, v):
//   apex v0 (shares v1 v2): u <= 1, v <= 1, u + v >= 1      apex v1 (shares v0 v2): u <= 0, v <= 1, u + v >= 0
//   apex v2 (shares v0 v1): u <= 1, v <= 0, u + v >= 0      -- i.e. u <= pu, v <= pv, u + v >= ps with (pu, pv, ps) in {0, 1}
struct UnitRec { float v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz, pu, pv, ps; };
// Branch-free: each triangle's three conditions and the t >= 0 condition (scaled by kx = 3 del / delx so that it shares the
// threshold) go through one minimum; when the orientation is uncertain the threshold becomes -infinity (everything is kept).
RLPT_HD unsigned unit_candidates(const UnitRec& r, float ox, float oy, float oz, float a0, float a1, float a2, float del, float kx) {
    const float bx = ox - r.v0x, by = oy - r.v0y, bz = oz - r.v0z;
    const float cx = a1 * bz - a2 * by, cy = a2 * bx - a0 * bz, cz = a0 * by - a1 * bx;
    const float D = a0 * r.nx + a1 * r.ny + a2 * r.nz;
    const float X = bx * r.nx + by * r.ny + bz * r.nz;
    const float Y = r.e2x * cx + r.e2y * cy + r.e2z * cz;
    const float Z = -(r.e1x * cx + r.e1y * cy + r.e1z * cz);
    const float sg = D < 0.f ? -1.f : 1.f;
    const float Ds = fabsf(D), Xk = X * (sg * kx), Ys = Y * sg, Zs = Z * sg, S = Ys + Zs;
    const float thr = Ds > del ? -3.f * del : -3.0e38f;     // NaN compares false: kept
    const float mA = fminf(fminf(Ys, Zs), fminf(Ds - S, Xk));
    const float mB = fminf(fminf(r.pu * Ds - Ys, r.pv * Ds - Zs), fminf(S - r.ps * Ds, Xk));
    return (!(mA < thr) ? 1u : 0u) | (!(mB < thr) ? 2u : 0u);
}

// ---------------------------------------------------------------- hemisphere helpers (G/utils/hemisphere_helpers.cu)
// create_normal_coordinate_system (:31-44); evaluated once per surface at upload, kept beside the triangle.
RLPT_HD void tangent_frame(f3 n, f3& T, f3& B) {
    f3 t = fabsf(n.x) > fabsf(n.y) ? f3{ n.z, 0.f, -n.x } : f3{ 0.f, -n.z, n.y };
    float inv = 1.f / sqrtf(t.x * t.x + t.y * t.y + t.z * t.z);
    T = f3{ t.x * inv, t.y * inv, t.z * inv };
    B = f3{ n.y * T.z - T.y * n.z, n.z * T.x - T.z * n.x, n.x * T.y - T.x * n.y };
}
// Shirley-Chiu concentric map of the unit square onto the hemisphere, the function map() computes (:134-226), in
// its four-quadrant form: radius r = max(|a|,|b|), angle from the quadrant; y_h = cos(theta) = 1 - r^2 and
// sin(theta) = r*sqrt(2 - r^2) (the reference takes acos(1-r^2) and then sin/cos of it).
RLPT_HD void square_to_hemisphere(float sx, float sy, float& xh, float& yh, float& zh) {
    const float a = 2.f * sx - 1.f, b = 2.f * sy - 1.f;
    const float q = 0.78539816339744830962f;   // pi/4
    // quadrant 0: a > |b|   r = a,  phi = q (b/a)        quadrant 1: b >= |a|  r = b,  phi = q (2 - a/b)
    // quadrant 2: -a > |b|  r = -a, phi = q (4 + b/a)    quadrant 3: otherwise r = -b, phi = q (6 - a/b), 0 at the centre
    // Written with selects and ONE division: the four-way branch ran each quadrant's lanes separately (6 of 32 lanes per pass,
    // 5 % of k_shade's instructions). Same roundings: k + (+-ratio) is the branchy form's 2 - a/b etc., and 0 + b/a is exact.
    const bool upper = a > -b, along_a = upper ? (a > b) : (a < b);
    const float den = along_a ? a : b, ratio = (along_a ? b : a) / den;
    const float k = upper ? (along_a ? 0.f : 2.f) : (along_a ? 4.f : 6.f);
    const float r = upper ? den : -den;
    float phi = q * (k + (along_a ? ratio : -ratio));
    if (!upper && !along_a && !(b != 0.f)) phi = 0.f;
    float s, c;
#if defined(__CUDA_ARCH__)
    sincosf(phi, &s, &c);
#else
    s = sinf(phi); c = cosf(phi);
#endif
    float st = r * sqrtf(2.f - r * r);
    xh = st * c; yh = 1.f - r * r; zh = st * s;
}
// convert_grid_pos_to_direction (:96-105): the matrix is a rotation (T,N,B) plus the volume position, and the
// position cancels in normalize(world - position); so the direction is the rotated hemisphere point.
RLPT_HD f3 grid_to_direction(float gx, float gy, f3 T, f3 N, f3 B) {
    float xh, yh, zh;
    square_to_hemisphere(gx * (1.f / 12.f), gy * (1.f / 12.f), xh, yh, zh);
    f3 w = { T.x * xh + N.x * yh + B.x * zh, T.y * xh + N.y * yh + B.y * zh, T.z * xh + N.z * yh + B.z * zh };
    float inv = 1.f / sqrtf(w.x * w.x + w.y * w.y + w.z * w.z);
    return f3{ w.x * inv, w.y * inv, w.z * inv };
}
// cos(theta) of a cell centre: dot(direction, N) = y_h = 1 - r^2 with r = max(|2x-1|,|2y-1|):
// it depends only on the cell, not on the volume.
RLPT_HD float cell_centre_cos(int cell) {
    float a = 2.f * (((float)(cell / GRID) + 0.5f) * (1.f / 12.f)) - 1.f, b = 2.f * (((float)(cell % GRID) + 0.5f) * (1.f / 12.f)) - 1.f;
    float r = fmaxf(fabsf(a), fabsf(b));
    return 1.f - r * r;
}
// sample_random_direction_around_intersection + uniform_hemisphere_sample (:8-25,67-93): cos(theta) = r1
RLPT_HD f3 uniform_hemisphere(float r1, float r2, f3 T, f3 N, f3 B) {
    float st = sqrtf(1.f - r1 * r1), phi = 6.28318530717958647692f * r2, s, c;
#if defined(__CUDA_ARCH__)
    sincosf(phi, &s, &c);
#else
    s = sinf(phi); c = cosf(phi);
#endif
    float x = st * c, z = st * s;
    return f3{ x * B.x + r1 * N.x + z * T.x, x * B.y + r1 * N.y + z * T.y, x * B.z + r1 * N.z + z * T.z };
}

// ---------------------------------------------------------------- radiance-volume CDF sampling
// RadianceVolume::sample_direction_from_radiance_distribution (G/radiance_volumes/radiance_volume.cu:192-244): bin 0 when
// r <= cdf[0], otherwise the bin with cdf[k-1] <= r < cdf[k]. Deliberate deviation: when r lies past the
// last bin the reference returns a zero direction and pdf 0 (a NaN sample); here it is clamped to the last bin of
// non-zero width. `load(k)` reads cdf[k] of the volume.
template <class Load>
RLPT_HD int sample_sector(Load load, float r, float& pdf) {
    float c0 = load(0);
    if (r <= c0) { pdf = RHO * (c0 / GRID_RHO); return 0; }
    int lo = 1, hi = CELLS;            // first k in [1,144) with cdf[k] > r
    while (lo < hi) { int mid = (lo + hi) >> 1; if (load(mid) > r) hi = mid; else lo = mid + 1; }
    int k = lo;
    if (k >= CELLS) { k = CELLS - 1; while (k > 0 && !(load(k) - load(k - 1) > 0.f)) --k; }
    float hi_v = load(k), lo_v = k > 0 ? load(k - 1) : 0.f;
    pdf = RHO * ((hi_v - lo_v) / GRID_RHO);
    return k;
}

// Two-level form of sample_sector with the same result: `rows` = cdf[12 j + 11] (12 values, written next to the CDF by the
// merge kernel) selects the grid row, then the row's 12 entries select the cell -- two rounds of independent 16-byte
// loads instead of a chain of ~10 dependent 4-byte probes. row(j, out[12]) loads cdf[12 j .. 12 j + 11].
template <class LoadRows, class LoadRow, class Load>
RLPT_HD int sample_sector_2level(LoadRows load_rows, LoadRow load_row, Load load, float r, float& pdf) {
    float e[12]; load_rows(e);
    int j = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 11; ++i) j += (e[i] <= r) ? 1 : 0;           // first row whose end exceeds r (monotone), capped at the last row
    float c[12]; load_row(j, c);
    if (j == 0 && r <= c[0]) { pdf = RHO * (c[0] / GRID_RHO); return 0; }
    int n = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 12; ++i) n += (c[i] <= r) ? 1 : 0;           // entries not exceeding r
    if (n < 12) {
        // the two CDF entries around the chosen cell are re-read by index (the row's line is in L1) rather than selected out
        // of the 24 registers with a chain of conditional moves
        const int k = 12 * j + n;
        const float hi_v = load(k), lo_v = k > 0 ? load(k - 1) : 0.f;
        pdf = RHO * ((hi_v - lo_v) / GRID_RHO);
        return k;
    }
    return sample_sector(load, r, pdf);                              // r at or past the end of the table: the clamping path
}

// ---------------------------------------------------------------- nearest radiance volume
// RadianceMap::find_closest_radiance_volume_iterative (G/radiance_volumes/radiance_map.cu:150-203) over the reference's
// own kd-tree, re-laid-out: inner nodes only, one float4 each (split, left, right, dim) with leaf children encoded
// in the child word as (KD_LEAF | volume); a leaf visit reads vol_posn[volume] = (position, normal class).
// Visit order, strict-< tie rule, the "within" test in double and the distance rounding are the reference's.
constexpr uint32_t KD_LEAF = 0x80000000u;
RLPT_HD float kd_distance(float px, float py, float pz, float qx, float qy, float qz) {
    float dx = RLPT_SUB(qx, px), dy = RLPT_SUB(qy, py), dz = RLPT_SUB(qz, pz);
    return RLPT_SQRT(RLPT_FMA(dz, dz, RLPT_FMA(dx, dx, RLPT_MUL(dy, dy))));
}
// `within_abs` is the largest float f with (double)f*f < (double)MAX_DIST, so that fabsf(delta) <= f is exactly the
// reference's `pow(delta, 2) < max_dist` (evaluated in double there, radiance_map.cu:185,193) without FP64 in the loop.
template <class LoadInner, class LoadVol>
RLPT_HD int kd_find(LoadInner load_inner, LoadVol load_vol, uint32_t root, float root_px, float root_py, float root_pz,
                    float px, float py, float pz, int normal_class, float within_abs) {
    uint32_t stack[32]; int top = 0;
    int best = 0; float best_d = kd_distance(px, py, pz, root_px, root_py, root_pz);
    uint32_t cur = root; bool have = true;
    while (have) {
        if (cur & KD_LEAF) {
            int vol = (int)(cur & ~KD_LEAF);
            float vx, vy, vz; int cls; load_vol(vol, vx, vy, vz, cls);
            float d = kd_distance(px, py, pz, vx, vy, vz);
            if (cls == normal_class && d < best_d) { best = vol; best_d = d; }
            if (top > 0) cur = stack[--top]; else have = false;
        } else {
            float split; uint32_t left, right; int dim; load_inner(cur, split, left, right, dim);
            float c = dim == 0 ? px : (dim == 1 ? py : pz);
            float delta = RLPT_SUB(c, split);
            bool within = fabsf(delta) <= within_abs;
            uint32_t near_c = (c < split) ? left : right, far_c = (c < split) ? right : left;
            if (within && top < 32) stack[top++] = far_c;
            cur = near_c;
        }
    }
    return best;
}

// Candidate-cell front end of the nearest-volume search (DESIGN.md "Nearest volume"). The kd search above visits every
// leaf whose position differs from the query by at most within_abs in every coordinate (a far child is entered when
// the query is within within_abs of the split plane, and the split lies between query and leaf), keeps the closest
// same-normal leaf with strict <, and starts from (volume 0, |query - root.position|). Hence: if the closest same-normal
// volume N among ALL volumes has distance r <= accept_r (accept_r < within_abs), the ball of radius r lies inside the
// visited box, N is the kd search's winner among leaves, and the kd answer is r < d0 ? N : 0.
// The closest same-normal volume is found from a precomputed list: a fine uniform grid (cells a fraction of the volume
// spacing) over the surfaces; for each (cell, normal class) the host lists every volume that can be the closest one
// within accept_r of some point of the cell (rlpt_radiance_host.cpp, host_build_vcells) -- a handful instead of the whole
// neighbourhood. (cell, class) -> list goes through an open-addressing hash table, one 16-byte entry per pair.
// Anything else -- no entry, no candidate within accept_r, two candidates whose rounded distances tie (the kd answer then
// depends on visit order), a query outside the grid -- returns -1 and the caller runs the kd search itself.
// Candidates are ranked on the squared distance (the argument of the reference's sqrt, same rounding sequence); only
// the winner and the runner-up are square-rooted, to detect ties after rounding.
struct VCells { float ox, oy, oz, inv_h, accept_r; int nx, ny, nz; uint32_t mask; };
RLPT_HD float grid_coord(float x, float o, float inv_h) { return floorf(RLPT_MUL(RLPT_SUB(x, o), inv_h)); }
RLPT_HD float kd_distance2(float px, float py, float pz, float qx, float qy, float qz) {
    float dx = RLPT_SUB(qx, px), dy = RLPT_SUB(qy, py), dz = RLPT_SUB(qz, pz);
    return RLPT_FMA(dz, dz, RLPT_FMA(dx, dx, RLPT_MUL(dy, dy)));
}
RLPT_HD uint32_t vcell_hash(uint32_t cell, uint32_t cls) {
    uint32_t h = (cell * 0x9E3779B1u) ^ (cls * 0x85EBCA77u);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
// load_entry(slot, cell, cls, start, n4) reads one table slot; load_cand(i, x, y, z, vol) one candidate (vol = -1: padding)
template <class LoadEntry, class LoadCand>
RLPT_HD int vcell_find(const VCells& g, LoadEntry load_entry, LoadCand load_cand, float px, float py, float pz, int normal_class, float d0) {
    float ux = grid_coord(px, g.ox, g.inv_h), uy = grid_coord(py, g.oy, g.inv_h), uz = grid_coord(pz, g.oz, g.inv_h);
    if (!(ux >= 0.f && uy >= 0.f && uz >= 0.f && ux < (float)g.nx && uy < (float)g.ny && uz < (float)g.nz)) return -1;
    const int cell = ((int)uz * g.ny + (int)uy) * g.nx + (int)ux;
    uint32_t slot = vcell_hash((uint32_t)cell, (uint32_t)normal_class) & g.mask;
    int start = 0, n4 = 0;
    for (int probe = 0; ; ++probe) {
        int ec, ek; load_entry(slot, ec, ek, start, n4);
        if (ec == cell && ek == normal_class) break;
        if (ec < 0 || probe >= 32) return -1;
        slot = (slot + 1) & g.mask;
    }
    float best2 = 3.0e38f, second2 = 3.0e38f; int best = -1;
    for (int i = 0; i < n4; ++i) {
        float vx[4], vy[4], vz[4]; int vol[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) load_cand(4 * (start + i) + k, vx[k], vy[k], vz[k], vol[k]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) {
            float d2 = kd_distance2(px, py, pz, vx[k], vy[k], vz[k]);
            second2 = fminf(second2, fmaxf(d2, best2));
            if (d2 < best2) { best2 = d2; best = vol[k]; }
        }
    }
    if (best < 0) return -1;
    const float best_d = RLPT_SQRT(best2);
    if (!(best_d <= g.accept_r) || RLPT_SQRT(second2) == best_d) return -1;
    return best_d < d0 ? best : 0;
}

// Second level, for the queries the first cannot decide because no same-class volume lies within accept_r: the kd search
// reaches leaf l from p exactly when fl(p_k - lo_k) >= -within_abs and fl(p_k - hi_k) <= within_abs in every dimension,
// [lo, hi] being l's kd cell (rlpt_radiance_host.cpp derives this from kd_find's far-child rule). The list of the query's
// (cell, class) holds every volume that can pass this test from the cell, so the closest one that does is the search's
// winner; an empty visited set leaves volume 0. Ties after rounding depend on the visit order: -1, the caller runs kd_find.
// load_cand(i, x, y, z, vol, lo[3], hi[3]).
template <class LoadEntry, class LoadCand>
RLPT_HD int vext_find(const VCells& g, uint32_t xmask, LoadEntry load_entry, LoadCand load_cand, float px, float py, float pz, int normal_class, float d0, float within_abs) {
    float ux = grid_coord(px, g.ox, g.inv_h), uy = grid_coord(py, g.oy, g.inv_h), uz = grid_coord(pz, g.oz, g.inv_h);
    if (!(ux >= 0.f && uy >= 0.f && uz >= 0.f && ux < (float)g.nx && uy < (float)g.ny && uz < (float)g.nz)) return -1;
    const int cell = ((int)uz * g.ny + (int)uy) * g.nx + (int)ux;
    uint32_t slot = vcell_hash((uint32_t)cell, (uint32_t)normal_class) & xmask;
    int start = 0, n = 0;
    for (int probe = 0; ; ++probe) {
        int ec, ek; load_entry(slot, ec, ek, start, n);
        if (ec == cell && ek == normal_class) break;
        if (ec < 0 || probe >= 32) return -1;
        slot = (slot + 1) & xmask;
    }
    float best2 = 3.0e38f, second2 = 3.0e38f; int best = -1;
    for (int i = 0; i < n; ++i) {
        float vx, vy, vz, lo[3], hi[3]; int vol; load_cand(start + i, vx, vy, vz, vol, lo, hi);
        const bool visited = RLPT_SUB(px, lo[0]) >= -within_abs && RLPT_SUB(px, hi[0]) <= within_abs && RLPT_SUB(py, lo[1]) >= -within_abs &&
                             RLPT_SUB(py, hi[1]) <= within_abs && RLPT_SUB(pz, lo[2]) >= -within_abs && RLPT_SUB(pz, hi[2]) <= within_abs;
        const float d2 = kd_distance2(px, py, pz, vx, vy, vz);
        if (visited) { second2 = fminf(second2, fmaxf(d2, best2)); if (d2 < best2) { best2 = d2; best = vol; } }
    }
    if (best < 0) return 0;
    const float best_d = RLPT_SQRT(best2);
    if (RLPT_SQRT(second2) == best_d) return -1;
    return best_d < d0 ? best : 0;
}

}  // namespace rlpt
