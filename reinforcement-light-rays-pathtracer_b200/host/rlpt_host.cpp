// rlpt_host.cpp -- implementation of the C++ host mirror (rlpt_host.h). Host-only code; everything that computes
// pixels or Q-values is a call into librlpt.so through include/rlpt.h. Compiled with -ffp-contract=off so the scene
// arithmetic rounds exactly as the reference's host code does (tests compare the loaders with the reference's output).
#include "rlpt_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace rlpt_host {

// ------------------------------------------------------------------------------------------------ small math
static inline vec4 scale4(vec4 v, float s) { return vec4(v.x * s, v.y * s, v.z * s, v.w * s); }
static inline float luminance_of(vec3 c) {                     // 0.5 * (max + min): material.cu:4-14, area_light.cu:13-21
    float mx = std::max(c.z, std::max(c.x, c.y)), mn = std::min(c.z, std::min(c.x, c.y));
    return 0.5f * (mx + mn);
}
Material::Material(vec3 c) : diffuse_c(c), luminance(luminance_of(c)) {}
AreaLight::AreaLight(vec4 a, vec4 b, vec4 c, vec3 p) : Triangle(a, b, c), diffuse_p(p), luminance(luminance_of(p)) {}

void Triangle::compute_and_set_normal() {                       // triangle.cu:67-76: normalize(cross(e2, e1))
    float e1[3] = { v1.x - v0.x, v1.y - v0.y, v1.z - v0.z }, e2[3] = { v2.x - v0.x, v2.y - v0.y, v2.z - v0.z };
    float c[3] = { e2[1] * e1[2] - e1[1] * e2[2], e2[2] * e1[0] - e1[2] * e2[0], e2[0] * e1[1] - e1[0] * e2[1] };
    float inv = 1.f / std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
    normal = vec4(c[0] * inv, c[1] * inv, c[2] * inv, 1.f);
}
float Triangle::compute_area() const {                          // triangle.cu:15-26
    float a[3] = { v1.x - v0.x, v1.y - v0.y, v1.z - v0.z }, b[3] = { v2.x - v0.x, v2.y - v0.y, v2.z - v0.z };
    float la = std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]), lb = std::sqrt((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
    float e = la * lb, cos_theta = ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) / e;
    float sin_theta = (float)std::sqrt(1 - std::pow((double)cos_theta, 2));
    return 0.5f * e * sin_theta;
}

// v * (2/l) - 1, then x and y negated: the placement both the Cornell builder (cornell_box_scene.cu:163-199) and
// build_area_lights (object_importer.cu:273-313) apply
static vec4 place(vec4 v, float s) {
    vec4 r = scale4(v, s);
    r = vec4(r.x - 1.f, r.y - 1.f, r.z - 1.f, r.w - 1.f);
    return vec4(r.x * -1.f, r.y * -1.f, r.z, 1.f);
}
static void push9(std::vector<float>& out, const vec4& a, const vec4& b, const vec4& c) {
    const float v[9] = { a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z };
    out.insert(out.end(), v, v + 9);
}
template <class T> static void place_all(std::vector<T>& tris, float s, std::vector<float>& vertices) {
    for (T& t : tris) { t.v0 = place(t.v0, s); t.v1 = place(t.v1, s); t.v2 = place(t.v2, s); push9(vertices, t.v0, t.v1, t.v2); t.compute_and_set_normal(); }
}

// ------------------------------------------------------------------------------------------------ Cornell box
void get_cornell_shapes(std::vector<Surface>& S, std::vector<AreaLight>& lights, std::vector<float>& vertices) {
    const Material blue(vec3(0.15f, 0.15f, 0.75f)), white(vec3(0.75f)), red(vec3(0.75f, 0.15f, 0.15f)), green(vec3(0.15f, 0.75f, 0.15f)),
        yellow(vec3(0.75f, 0.75f, 0.15f)), cyan(vec3(0.15f, 0.75f, 0.75f));
    const float l = 555;
    vec4 A(l, 0, 0, 1), B(0, 0, 0, 1), C(l, 0, l, 1), D(0, 0, l, 1), E(l, l, 0, 1), F(0, l, 0, 1), G(l, l, l, 1), H(0, l, l, 1);
    const vec4 I(l / 3, l, (2 * l) / 3, 1), J((2 * l) / 3, l, (2 * l) / 3, 1), K(l / 3, l, l / 3, 1), L((2 * l) / 3, l, l / 3, 1);
    auto tri = [&](vec4 a, vec4 b, vec4 c, const Material& m) { S.push_back(Surface(a, b, c, m)); };
    tri(C, B, A, green); tri(C, D, B, green);                                   // floor
    tri(A, E, C, white); tri(C, E, G, white);                                   // left wall
    tri(F, B, D, white); tri(H, F, D, white);                                   // right wall
    tri(F, H, I, cyan); tri(F, I, K, cyan); tri(F, K, E, cyan); tri(K, L, E, cyan);   // ceiling around the light
    tri(L, G, E, cyan); tri(L, J, G, cyan); tri(I, G, J, cyan); tri(H, G, I, cyan);
    const vec3 diffuse_p(14.f * 0.9f, 14.f * 0.9f, 14.f * 0.9f);
    lights.push_back(AreaLight(K, I, J, diffuse_p)); lights.push_back(AreaLight(K, J, L, diffuse_p));
    tri(G, D, C, yellow); tri(G, H, D, yellow);                                 // back wall
    auto block = [&](const Material& m) {
        tri(E, B, A, m); tri(E, F, B, m); tri(F, D, B, m); tri(F, H, D, m); tri(H, C, D, m); tri(H, G, C, m);
        tri(G, E, C, m); tri(E, A, C, m); tri(G, F, E, m); tri(G, H, F, m);
    };
    A = vec4(240, 0, 234, 1); B = vec4(80, 0, 185, 1); C = vec4(190, 0, 392, 1); D = vec4(32, 0, 345, 1);
    E = vec4(240, 165, 234, 1); F = vec4(80, 165, 185, 1); G = vec4(190, 165, 392, 1); H = vec4(32, 165, 345, 1);
    block(blue);                                                                // short block
    A = vec4(443, 0, 247, 1); B = vec4(285, 0, 296, 1); C = vec4(492, 0, 406, 1); D = vec4(334, 0, 456, 1);
    E = vec4(443, 330, 247, 1); F = vec4(285, 330, 296, 1); G = vec4(492, 330, 406, 1); H = vec4(334, 330, 456, 1);
    block(red);                                                                 // tall block
    place_all(S, 2 / l, vertices);
    place_all(lights, 2 / l, vertices);
}

// ------------------------------------------------------------------------------------------------ .obj import
static void split_nonempty(std::vector<std::string>& out, std::string s, const std::string& delim) {    // object_importer.cu:188-206
    size_t pos;
    while ((pos = s.find(delim)) != std::string::npos) { std::string tok = s.substr(0, pos); if (!tok.empty()) out.push_back(tok); s.erase(0, pos + delim.size()); }
    out.push_back(s);
}

static void preset_lights(const ImportPreset& p, std::vector<AreaLight>& lights, std::vector<float>& vertices) {   // object_importer.cu:209-314
    const vec3 power(8.f * 1.f, 8.f * 1.f, 8.f * 1.f);
    const float l = 2.f;
    auto quad = [&](vec4 I, vec4 J, vec4 K, vec4 L) { lights.push_back(AreaLight(K, I, J, power)); lights.push_back(AreaLight(K, J, L, power)); };
    switch (p.lights) {
    case ImportPreset::DOOR_ROOM:
        quad(vec4((6.3f * l) / 8, (l * 6.f) / 8, 1.499f * l, 1), vec4((6.3f * l) / 8, 0, 1.499f * l, 1),
             vec4((2.58f * l) / 8, (l * 6.f) / 8, 1.499f * l, 1), vec4((2.58f * l) / 8, 0, 1.499f * l, 1)); break;
    case ImportPreset::SIMPLE_CLOSED_ROOM:
        quad(vec4(l - 0.001f, (l * 4.f) / 8, 1.f * l, 1.f), vec4(l - 0.001f, (l * 1.f) / 8, 1.f * l, 1.f),
             vec4(l - 0.001f, (l * 4.f) / 8, 0.5f * l, 1.f), vec4(l - 0.001f, (l * 1.f) / 8, 0.5f * l, 1.f)); break;
    case ImportPreset::SIMPLE_ROOM:
        quad(vec4(l - 0.001f, (l * 6.f) / 8, 0.5f * l, 1.f), vec4(l - 0.001f, (l * 3.f) / 8, 0.5f * l, 1.f),
             vec4(l - 0.001f, (l * 6.f) / 8, 0.25f * l, 1.f), vec4(l - 0.001f, (l * 3.f) / 8, 0.25f * l, 1.f)); break;
    case ImportPreset::ARCHWAY: {
        quad(vec4(l + 1.99f, l, (float)(2.5 * l), 1.f), vec4(l + 1.99f, (l * 4.f) / 8, 2.5f * l, 1.f), vec4(l + 1.99f, l, 2.f * l, 1.f), vec4(l + 1.99f, (l * 4.f) / 8, 2.f * l, 1.f));
        const vec4 M(l - 1.99f, l, 2.5f * l, 1.f), N(l - 1.99f, (l * 4.f) / 8, 2.5f * l, 1.f), O(l - 1.99f, l, 2.0f * l, 1.f), P(l - 1.99f, (l * 4.f) / 8, 2.0f * l, 1.f);
        lights.push_back(AreaLight(O, M, N, power)); lights.push_back(AreaLight(O, N, P, power));
        const vec4 Q(l - 0.5f, l, 2.99f * l, 1.f), R(l - 0.5f, l * 0.5f, 2.99f * l, 1.f), S(l + 0.5f, l, 2.99f * l, 1.f), T(l + 0.5f, l * 0.5f, 2.99f * l, 1.f);
        lights.push_back(AreaLight(S, Q, R, power)); lights.push_back(AreaLight(S, R, T, power));
        break; }
    case ImportPreset::NONE: break;
    }
    place_all(lights, 2 / l, vertices);
}

static void build_from_obj(std::vector<Surface>& surfaces, std::vector<AreaLight>& lights, std::vector<float>& vertices,
                           const std::vector<int>& idx /* 3 per face */, const std::vector<vec3>& verts, bool lights_in_obj, const ImportPreset& preset) {
    // object_importer.cu:92-185 / :317-412. min/max start at 0, not at the first vertex (:96-97).
    float max_pos[3] = { 0.f, 0.f, 0.f }, min_pos[3] = { 0.f, 0.f, 0.f };
    for (const vec3& v : verts) {
        const float c[3] = { v.x, v.y, v.z };
        for (int k = 0; k < 3; ++k) { if (c[k] > max_pos[k]) max_pos[k] = c[k]; if (c[k] < min_pos[k]) min_pos[k] = c[k]; }
    }
    float max_difference = 0.f;
    for (int k = 0; k < 3; ++k) max_difference = std::max(max_difference, std::fabs(max_pos[k] - min_pos[k]));
    const float scale = preset.normalise && max_difference > 0.f ? 2.f / max_difference : 2.f;
    const float dist[3] = { -1.f - (min_pos[0] * scale), -1.f - (min_pos[1] * scale), -1.f - (min_pos[2] * scale) };
    auto xf = [&](const vec3& v) { return vec4((v.x * scale + dist[0]) * -1.f, (v.y * scale + dist[1]) * -1.f, v.z * scale + dist[2], 1.f); };
    const int n_faces = (int)idx.size() / 3;
    for (int i = 0; i < n_faces; ++i) {
        const int a = idx[3 * i] - 1, b = idx[3 * i + 1] - 1, c = idx[3 * i + 2] - 1;
        if (a < 0 || b < 0 || c < 0 || a >= (int)verts.size() || b >= (int)verts.size() || c >= (int)verts.size()) continue;   // the reference would read out of bounds
        const vec4 v1 = xf(verts[a]), v2 = xf(verts[b]), v3 = xf(verts[c]);
        if (lights_in_obj) {
            if ((i > 23 && i < 36) || (i > 50 && i < 63)) {
                AreaLight al(v1, v3, v2, vec3(12.f * 1.f)); al.compute_and_set_normal(); lights.push_back(al);
            } else {
                Material m(vec3(0.9f));
                if (i >= 0 && i <= 7) m = Material(vec3(0.1f)); else if (i > 133 && i < 142) m = Material(vec3(0.75f, 0.15f, 0.15f));
                Surface s(v1, v3, v2, m); s.compute_and_set_normal(); surfaces.push_back(s);
            }
        } else {
            Material m(vec3(0.75f));
            if (preset.colours == ImportPreset::COMMITTED) { if (i > 80) m = Material(vec3(0.75f, 0.15f, 0.15f)); if (11 < i && i < 24) m = Material(vec3(0.15f, 0.15f, 0.75f)); }
            else if (preset.colours == ImportPreset::DOOR_ROOM_COLOURS) { if (i > 23 && i < 36) m = Material(vec3(0.75f, 0.15f, 0.15f)); if (11 < i && i < 24) m = Material(vec3(0.15f, 0.15f, 0.75f)); }
            Surface s(v1, v3, v2, m); s.compute_and_set_normal(); surfaces.push_back(s);          // winding (v1, v3, v2): :166
        }
        push9(vertices, v1, v2, v3);
    }
    if (!lights_in_obj) preset_lights(preset, lights, vertices);
}

bool load_scene(const char* path, std::vector<Surface>& surfaces, std::vector<AreaLight>& lights, std::vector<float>& vertices, bool lights_in_obj, const ImportPreset& preset) {
    FILE* file = fopen(path, "r");
    if (!file) { printf("File %s could not be opened!\n", path); return false; }
    std::vector<vec3> verts; std::vector<int> idx;
    char header[128];
    while (fscanf(file, "%127s", header) != EOF) {
        if (strcmp(header, "v") == 0) { vec3 v; if (fscanf(file, "%f %f %f\n", &v.x, &v.y, &v.z) == 3) verts.push_back(v); }
        else if (strcmp(header, "f") == 0) {
            char line[256];
            if (!fgets(line, sizeof line, file)) continue;
            std::vector<std::string> args; split_nonempty(args, line, " ");
            std::vector<int> face;
            for (const std::string& a : args) {                                   // "i/j/k" or plain "i": the leading integer is the vertex
                if (a.empty() || a == "\n" || a == "\r\n") continue;
                try { face.push_back(std::stoi(a.substr(0, a.find('/')), nullptr, 10)); } catch (...) {}
            }
            for (size_t k = 1; k + 1 < face.size(); ++k) { idx.push_back(face[0]); idx.push_back(face[k]); idx.push_back(face[k + 1]); }   // fan triangulation
        }
    }
    fclose(file);
    build_from_obj(surfaces, lights, vertices, idx, verts, lights_in_obj, preset);
    return true;
}

// ------------------------------------------------------------------------------------------------ Scene
Scene::~Scene() { delete[] surfaces; delete[] area_lights; delete[] vertices; }
void Scene::adopt(std::vector<Surface>& s, std::vector<AreaLight>& l, std::vector<float>& v) {
    delete[] surfaces; delete[] area_lights; delete[] vertices;
    surfaces_count = (int)s.size(); area_light_count = (int)l.size(); vertices_count = (int)v.size();
    surfaces = new Surface[std::max(surfaces_count, 1)]; area_lights = new AreaLight[std::max(area_light_count, 1)]; vertices = new float[std::max(vertices_count, 1)];
    std::copy(s.begin(), s.end(), surfaces); std::copy(l.begin(), l.end(), area_lights); std::copy(v.begin(), v.end(), vertices);
}
void Scene::load_cornell_box_scene() { std::vector<Surface> s; std::vector<AreaLight> l; std::vector<float> v; get_cornell_shapes(s, l, v); adopt(s, l, v); }
bool Scene::load_custom_scene(const char* filename, bool lights_in_obj, const ImportPreset& preset, const char* saved_radiance_volumes) {
    std::vector<Surface> s; std::vector<AreaLight> l; std::vector<float> v;
    bool ok = load_scene(filename, s, l, v, lights_in_obj, preset);
    if (saved_radiance_volumes) read_radiance_volumes_to_surfaces(saved_radiance_volumes, s);      // RENDER_SAVED_RADIANCE_VOLUMES (scene.cu:41-46); `vertices` stays the .obj's
    adopt(s, l, v);
    return ok;
}
void Scene::save_vertices_to_file(const char* path) const {
    std::ofstream f(path);
    if (!f.is_open()) { printf("Unable to save the vertices.\n"); return; }
    auto line = [&](const Triangle& t) { f << t.v0.x << " " << t.v0.y << " " << t.v0.z << " " << t.v1.x << " " << t.v1.y << " " << t.v1.z << " " << t.v2.x << " " << t.v2.y << " " << t.v2.z << "\n"; };
    for (int i = 0; i < surfaces_count; ++i) line(surfaces[i]);
    for (int i = 0; i < area_light_count; ++i) line(area_lights[i]);
}

// ------------------------------------------------------------------------------------------------ Camera (G/camera.cu:3-83)
static void mat_identity(float R[4][4]) { for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) R[c][r] = c == r ? 1.f : 0.f; }
static vec4 mat_mul(const float R[4][4], vec4 v) {              // column-major like glm: R[col][row]
    return vec4(R[0][0] * v.x + R[1][0] * v.y + R[2][0] * v.z + R[3][0] * v.w, R[0][1] * v.x + R[1][1] * v.y + R[2][1] * v.z + R[3][1] * v.w,
                R[0][2] * v.x + R[1][2] * v.y + R[2][2] * v.z + R[3][2] * v.w, R[0][3] * v.x + R[1][3] * v.y + R[2][3] * v.z + R[3][3] * v.w);
}
Camera::Camera(vec4 p) : position(p) { mat_identity(R); }
static void set_col(float R[4][4], int c, float a, float b, float d, float e) { R[c][0] = a; R[c][1] = b; R[c][2] = d; R[c][3] = e; }
void Camera::rotate_left(float y) { yaw_y += y; set_col(R, 0, std::cos(y), 0, std::sin(y), 0); set_col(R, 2, -std::sin(y), 0, std::cos(y), 0); position = mat_mul(R, position); }
void Camera::rotate_right(float y) { yaw_y -= y; set_col(R, 0, std::cos(-y), 0, std::sin(-y), 0); set_col(R, 2, -std::sin(-y), 0, std::cos(-y), 0); position = mat_mul(R, position); }
void Camera::rotate_up(float x) { yaw_x -= x; set_col(R, 0, 1.f, 0, 0, 0); set_col(R, 1, 0, std::cos(-x), -std::sin(-x), 0); set_col(R, 2, 0, std::sin(-x), std::cos(-x), 0); position = mat_mul(R, position); }
void Camera::rotate_down(float x) { yaw_x += x; set_col(R, 0, 1.f, 0, 0, 0); set_col(R, 1, 0, std::cos(x), -std::sin(x), 0); set_col(R, 2, 0, std::sin(x), std::cos(x), 0); position = mat_mul(R, position); }
// look_at(from, to) * (0,0,0,1) is `from` itself (camera.cu:58-83), so the move reduces to the translated position
void Camera::move_forwards(float d) { position = vec4(position.x - d * std::sin(yaw_y), position.y, position.z + d * std::cos(yaw_y), 1.f); }
void Camera::move_backwards(float d) { position = vec4(position.x + d * std::sin(yaw_y), position.y, position.z - d * std::cos(yaw_y), 1.f); }

// ------------------------------------------------------------------------------------------------ SDLScreen (headless)
void SDLScreen::PutPixelSDL(int x, int y, vec3 c) {
    if (x < 0 || x >= width || y < 0 || y >= height) { printf("apa\n"); return; }
    auto q = [](float v) { return (uint32_t)std::min(std::max(255 * v, 0.f), 255.f); };
    buffer[(size_t)y * width + x] = (128u << 24) + (q(c.x) << 16) + (q(c.y) << 8) + q(c.z);
}
bool SDLScreen::SDL_SaveImage(const char* filename) const {     // the file SDL_SaveBMP writes for an ARGB8888 surface (Images/render.bmp)
    FILE* f = fopen(filename, "wb");
    if (!f) return false;
    auto u16 = [&](uint16_t v) { fwrite(&v, 2, 1, f); }; auto u32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
    const uint32_t data = (uint32_t)width * height * 4, off = 14 + 108;
    fputc('B', f); fputc('M', f); u32(off + data); u16(0); u16(0); u32(off);
    u32(108); u32((uint32_t)width); u32((uint32_t)height); u16(1); u16(32); u32(3); u32(data); u32(0); u32(0); u32(0); u32(0);
    u32(0x00ff0000u); u32(0x0000ff00u); u32(0x000000ffu); u32(0xff000000u); u32(0x57696e20u);
    for (int i = 0; i < 12; ++i) u32(0);
    for (int y = height - 1; y >= 0; --y) fwrite(&buffer[(size_t)y * width], 4, (size_t)width, f);
    fclose(f);
    return true;
}

// ------------------------------------------------------------------------------------------------ Renderer
void Renderer::check(int status) const { if (status != RLPT_OK) throw RenderError{ status, rlpt_last_error() }; }
Renderer::Renderer(int device) { check(rlpt_ctx_create(device, &ctx_)); check(rlpt_config_get(ctx_, &cfg_)); }
Renderer::~Renderer() { rlpt_ctx_destroy(ctx_); }
void Renderer::apply_settings() { check(rlpt_config_set(ctx_, &cfg_)); }
void Renderer::upload(const Scene& s) {
    std::vector<float> sv((size_t)9 * s.surfaces_count), srgb((size_t)3 * s.surfaces_count), lv((size_t)9 * s.area_light_count), lrgb((size_t)3 * s.area_light_count);
    auto put = [](float* d, const Triangle& t) { const float v[9] = { t.v0.x, t.v0.y, t.v0.z, t.v1.x, t.v1.y, t.v1.z, t.v2.x, t.v2.y, t.v2.z }; std::copy(v, v + 9, d); };
    for (int i = 0; i < s.surfaces_count; ++i) { put(&sv[9 * (size_t)i], s.surfaces[i]); const vec3& c = s.surfaces[i].material.diffuse_c; srgb[3 * (size_t)i] = c.x; srgb[3 * (size_t)i + 1] = c.y; srgb[3 * (size_t)i + 2] = c.z; }
    for (int i = 0; i < s.area_light_count; ++i) { put(&lv[9 * (size_t)i], s.area_lights[i]); const vec3& c = s.area_lights[i].diffuse_p; lrgb[3 * (size_t)i] = c.x; lrgb[3 * (size_t)i + 1] = c.y; lrgb[3 * (size_t)i + 2] = c.z; }
    check(rlpt_scene_upload(ctx_, sv.data(), srgb.data(), s.surfaces_count, lv.data(), lrgb.data(), s.area_light_count));
}
void Renderer::set_camera(const Camera& c) { const float p[4] = { c.position.x, c.position.y, c.position.z, 1.f }; check(rlpt_camera_set(ctx_, p, c.yaw_y, c.yaw_x)); }
void Renderer::render_default(int frames) { check(rlpt_render_default(ctx_, frames)); }
void Renderer::render_sarsa(int frames) { check(rlpt_render_sarsa(ctx_, frames)); }
void Renderer::reset_frame() { check(rlpt_frame_reset(ctx_)); }
void Renderer::download(std::vector<float>& rgb) { rgb.resize((size_t)3 * cfg_.width * cfg_.height); check(rlpt_frame_download(ctx_, rgb.data())); }
void Renderer::present(SDLScreen& screen) {
    if (screen.width != cfg_.width || screen.height != cfg_.height) throw RenderError{ RLPT_ERR_ARG, "SDLScreen size differs from the render settings" };
    check(rlpt_frame_download_argb(ctx_, screen.buffer.data()));                // PutPixelSDL for every pixel, done on the device
}
rlpt_stats_t Renderer::stats() { rlpt_stats_t s; check(rlpt_stats(ctx_, &s)); return s; }
void Renderer::append_training_stats(const char* path) {
    rlpt_stats_t s = stats();
    // the reference prints int(total / pixels) (integer division, G/main.cu:327); kept
    float avg = s.paths > 0 ? (float)(long long)(s.path_length_sum / s.paths) : 0.f;
    std::ofstream f(path, std::ios::app);
    f << avg << " " << 0.0 << " " << (long long)s.zero_contribution_paths << "\n";
}

// ------------------------------------------------------------------------------------------------ RadianceMap
RadianceMap::RadianceMap(Renderer& r) : r_(r) {
    r_.check(rlpt_radiance_map_build(r_.ctx()));
    r_.check(rlpt_radiance_map_info(r_.ctx(), &radiance_volumes_count, &radiance_array_size));
}
void RadianceMap::update_radiance_volume_distributions() { r_.check(rlpt_radiance_map_update_distributions(r_.ctx())); }
void RadianceMap::save_q_vals_to_file(const char* path) { r_.check(rlpt_radiance_map_save_q(r_.ctx(), path)); }
void RadianceMap::load_q_vals_from_file(const char* path) { r_.check(rlpt_radiance_map_load_q(r_.ctx(), path)); }

bool read_hemisphere_locations_and_normals(const std::string& path, std::vector<vec3>& locations, std::vector<vec3>& normals) {
    std::ifstream in(path);
    if (!in.is_open()) { printf("Cannot read in hemisphere locations and normals.\n"); return false; }
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line); float v[6]; int n = 0;
        while (n < 6 && (ss >> v[n])) ++n;
        if (n < 6) continue;
        locations.push_back(vec3(v[0], v[1], v[2])); normals.push_back(vec3(v[3], v[4], v[5]));
    }
    return true;
}
void RadianceMap::save_selected_radiance_volumes_vals(std::string fpath) {
    const std::string read_in = fpath + "to_select.txt", write_out = fpath + "selected_sarsa.txt";
    std::remove(write_out.c_str());
    std::vector<vec3> loc, nrm;
    if (!read_hemisphere_locations_and_normals(read_in, loc, nrm) || loc.empty()) return;
    const int n = (int)loc.size(), nv = radiance_volumes_count;
    std::vector<float> pos(3 * (size_t)n), nn(3 * (size_t)n); std::vector<int> found(n);
    for (int i = 0; i < n; ++i) { pos[3 * i] = loc[i].x; pos[3 * i + 1] = loc[i].y; pos[3 * i + 2] = loc[i].z; nn[3 * i] = nrm[i].x; nn[3 * i + 1] = nrm[i].y; nn[3 * i + 2] = nrm[i].z; }
    r_.check(rlpt_radiance_map_find_closest(r_.ctx(), pos.data(), nn.data(), n, found.data()));
    std::vector<float> cdf((size_t)nv * RLPT_GRID_CELLS), vpos(3 * (size_t)nv), vnrm(3 * (size_t)nv);
    r_.check(rlpt_radiance_map_download(r_.ctx(), nullptr, cdf.data(), nullptr, nullptr, vpos.data(), vnrm.data(), nullptr));
    std::ofstream f(write_out, std::ios::app);
    if (!f.is_open()) { printf("Unable to save the Radiance Volume.\n"); return; }
    for (int i = 0; i < n; ++i) {                                 // RadianceVolume::write_volume_to_file (radiance_volume.cu:340-365)
        const int v = found[i];
        f << vpos[3 * v] << " " << vpos[3 * v + 1] << " " << vpos[3 * v + 2] << " " << vnrm[3 * v] << " " << vnrm[3 * v + 1] << " " << vnrm[3 * v + 2];
        // the reference converts the CDF to per-bin probabilities first (convert_radiance_volumes_distributions, G/main.cu:376)
        for (int k = 0; k < RLPT_GRID_CELLS; ++k) f << " " << (cdf[(size_t)v * RLPT_GRID_CELLS + k] - (k ? cdf[(size_t)v * RLPT_GRID_CELLS + k - 1] : 0.f));
        f << "\n";
    }
}

// ------------------------------------------------------------------------------------------------ saved radiance volumes as geometry
// Shirley-Chiu concentric map of the unit square onto the hemisphere around +y (hemisphere_helpers.cu:134-226). The square is
// cut into eight wedges by the axes and diagonals of (2x-1, 2y-1); wedge k starts at azimuth k pi/4, `r` is the distance from
// the centre in the wedge's leading coordinate, `s` runs across the wedge. cos(theta) = 1 - r^2. Intermediate types follow the
// reference: the wedge's start angle is a float, the azimuth is summed in double and rounded once.
void map(float x, float y, float& x_ret, float& y_ret, float& z_ret) {
    const float u = 2 * x - 1, v = 2 * y - 1;
    int wedge; float r, s;
    if (v > -u) {
        if (v < u) { r = u; if (v > 0) { wedge = 0; s = v; } else { wedge = 7; s = u + v; } }
        else       { r = v; if (u > 0) { wedge = 1; s = v - u; } else { wedge = 2; s = -u; } }
    } else {
        if (v > u) { r = -u; if (v > 0) { wedge = 3; s = -u - v; } else { wedge = 4; s = -v; } }
        else {
            r = -v;
            if (u > 0) { wedge = 6; s = u; }
            else if (v != 0) { wedge = 5; s = u - v; }
            else { x_ret = 0.f; y_ret = 1.f; z_ret = 0.f; return; }             // the centre of the square: straight up
        }
    }
    const float start = (float)((wedge * M_PI) / 4);
    const float theta = std::acos(1 - r * r);
    const float phi = (float)(start + (M_PI / 4) * (s / r));
    x_ret = std::sin(theta) * std::cos(phi);
    y_ret = std::cos(theta);
    z_ret = std::sin(theta) * std::sin(phi);
}
// create_normal_coordinate_system (hemisphere_helpers.cu:31-44)
static void normal_frame(vec3 n, vec3& t, vec3& b) {
    t = std::fabs(n.x) > std::fabs(n.y) ? vec3(n.z, 0.f, -n.x) : vec3(0.f, -n.z, n.y);
    const float inv = 1.f / std::sqrt((t.x * t.x + t.y * t.y) + t.z * t.z);
    t = vec3(t.x * inv, t.y * inv, t.z * inv);
    b = vec3(n.y * t.z - t.y * n.z, n.z * t.x - t.z * n.x, n.x * t.y - t.x * n.y);           // cross(n, t)
}
std::vector<std::vector<vec4>> SavedRadianceVolume::get_vertices() const {
    const float kDiameter = 0.15f;                                                 // DIAMETER, radiance_volumes_settings.h:11
    vec3 t, b; normal_frame(normal, t, b);
    // the volume's local -> world matrix has columns (T, N, B, position) (create_transformation_matrix, hemisphere_helpers.cu:48-63);
    // a mat4 * vec4 product sums its columns pairwise: (c0 x + c1 y) + (c2 z + c3 w)
    std::vector<std::vector<vec4>> grid(RLPT_GRID_RESOLUTION + 1, std::vector<vec4>(RLPT_GRID_RESOLUTION + 1));
    for (int gx = 0; gx <= RLPT_GRID_RESOLUTION; ++gx)
        for (int gy = 0; gy <= RLPT_GRID_RESOLUTION; ++gy) {
            float hx, hy, hz; map(gx / (float)RLPT_GRID_RESOLUTION, gy / (float)RLPT_GRID_RESOLUTION, hx, hy, hz);
            hx *= kDiameter; hy *= kDiameter; hz *= kDiameter;
            grid[gx][gy] = vec4((t.x * hx + normal.x * hy) + (b.x * hz + position.x * 1.f), (t.y * hx + normal.y * hy) + (b.y * hz + position.y * 1.f),
                                (t.z * hx + normal.z * hy) + (b.z * hz + position.z * 1.f), (0.f * hx + 0.f * hy) + (0.f * hz + 1.f * 1.f));
        }
    return grid;
}
void SavedRadianceVolume::build_surfaces(std::vector<Surface>& surfaces) const {
    float top = 0.f;
    for (int k = 0; k < RLPT_GRID_CELLS; ++k) if (top < radiance_distribution[k]) top = radiance_distribution[k];
    const std::vector<std::vector<vec4>> g = get_vertices();
    for (int gx = 0; gx < RLPT_GRID_RESOLUTION; ++gx)
        for (int gy = 0; gy < RLPT_GRID_RESOLUTION; ++gy) {
            const vec4 a = g[gx][gy], bq = g[gx + 1][gy], c = g[gx][gy + 1], d = g[gx + 1][gy + 1];
            const vec4 mid(((a.x + bq.x) + c.x + d.x) / 4.f, ((a.y + bq.y) + c.y + d.y) / 4.f, ((a.z + bq.z) + c.z + d.z) / 4.f, ((a.w + bq.w) + c.w + d.w) / 4.f);
            const float ratio = radiance_distribution[gx * RLPT_GRID_RESOLUTION + gy] / top;
            const Material colour(vec3(ratio, 1.f - ratio, 0.f));
            // both triangles of the quad face away from the volume's centre: normalize(mid - position) as a vec4 (w = 0)
            const float nx = mid.x - position.x, ny = mid.y - position.y, nz = mid.z - position.z, nw = mid.w - position.w;
            const float inv = 1.f / std::sqrt(((nx * nx + ny * ny) + nz * nz) + nw * nw);
            Surface s1(a, c, bq, colour), s2(bq, c, d, colour);
            s1.normal = s2.normal = vec4(nx * inv, ny * inv, nz * inv, nw * inv);
            surfaces.push_back(s1); surfaces.push_back(s2);
        }
}
bool read_radiance_volumes_from_file(const std::string& fname, std::vector<SavedRadianceVolume>& rvs) {
    std::ifstream in(fname);
    if (!in.is_open()) { printf("Could not read radiance volumes.\n"); return false; }
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line); std::string tok; std::vector<float> val;
        while (std::getline(ss, tok, ' ')) if (!tok.empty()) val.push_back(std::stof(tok));
        if (val.size() < 6 + (size_t)RLPT_GRID_CELLS) continue;                  // ragged line: skipped (the reference reads past the end of its vector)
        SavedRadianceVolume rv;
        rv.position = vec4(val[0], val[1], val[2], 1.f); rv.normal = vec3(val[3], val[4], val[5]);
        for (int k = 0; k < RLPT_GRID_CELLS; ++k) rv.radiance_distribution[k] = val[6 + k];
        rvs.push_back(rv);
    }
    return true;
}
bool read_radiance_volumes_to_surfaces(const std::string& fname, std::vector<Surface>& surfaces) {
    std::vector<SavedRadianceVolume> rvs;
    if (!read_radiance_volumes_from_file(fname, rvs)) return false;
    for (const SavedRadianceVolume& rv : rvs) rv.build_surfaces(surfaces);
    return true;
}

// ------------------------------------------------------------------------------------------------ Neural-Q drivers
static void upload_with_vertices(Renderer& r, const Scene& scene) {
    r.upload(scene);
    if (scene.vertices_count == 9 * (scene.surfaces_count + scene.area_light_count))       // the network's input order is Scene::vertices (G/main.cu:171-176)
        r.check(rlpt_dqn_set_vertices(r.ctx(), scene.vertices, scene.vertices_count));
}
NeuralQPathtracer::NeuralQPathtracer(unsigned int frames, int batch_size, SDLScreen& screen, Renderer& r, Scene& scene, Camera& camera, int, char**,
                                     const char* load_model, const char* save_model, const char* stats_file, const char* image) {
    upload_with_vertices(r, scene); r.set_camera(camera);
    bool loaded = false;
    if (load_model) { FILE* f = fopen(load_model, "r"); if (f) { fclose(f); r.check(rlpt_dqn_load_text(r.ctx(), load_model)); loaded = true; } }
    if (!loaded) r.check(rlpt_dqn_init(r.ctx(), 1984u));
    for (unsigned int f = 0; f < frames; ++f) {
        r.reset_frame(); r.check(rlpt_stats_reset(r.ctx()));
        r.check(rlpt_render_neuralq(r.ctx(), 1, batch_size));
        r.check(rlpt_neuralq_last_loss(r.ctx(), &last_loss));
        if (stats_file) {                                                      // "avg_path_length loss zero_contribution_paths" (neural_q_pathtracer.cu:578-583)
            rlpt_stats_t st = r.stats(); std::ofstream out(stats_file, std::ios::app);
            out << (st.paths > 0 ? st.path_length_sum / st.paths : 0.0) << " " << last_loss << " " << (long long)st.zero_contribution_paths << "\n";
        }
        r.present(screen); screen.SDL_Renderframe();
        if (save_model) r.check(rlpt_dqn_save_text(r.ctx(), save_model));    // after every frame (:191-196)
    }
    if (image) screen.SDL_SaveImage(image);
}
PretrainedPathtracer::PretrainedPathtracer(unsigned int frames, int, SDLScreen& screen, Renderer& r, Scene& scene, Camera& camera, int, char**, const char* model, const char* image) {
    FILE* f = model ? fopen(model, "r") : nullptr;
    if (!f) return;
    fclose(f);
    upload_with_vertices(r, scene); r.set_camera(camera);
    r.check(rlpt_dqn_load_text(r.ctx(), model));
    for (unsigned int fr = 0; fr < frames; ++fr) { r.reset_frame(); r.check(rlpt_render_pretrained(r.ctx(), 1)); r.present(screen); screen.SDL_Renderframe(); }
    if (image) screen.SDL_SaveImage(image);
    rendered = true;
}


// ---- NN_Q_Value_Trainer/Source/main.cu: the offline supervised trainer over the C ABI
static bool read_float_lines(const std::string& path, std::vector<std::vector<float>>& rows) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::string line;
    while (std::getline(in, line)) {
        std::vector<float> v; const char* sp = line.c_str(); char* end = nullptr;
        for (;;) { float f = strtof(sp, &end); if (end == sp) break; v.push_back(f); sp = end; }
        rows.push_back(v);
    }
    return true;
}
bool load_radiance_map_data(const std::string& path, std::vector<std::vector<float>>& data, int& action_count) {
    std::vector<std::vector<float>> rows;
    if (!read_float_lines(path, rows) || rows.empty() || rows[0].size() != 1) { printf("Radiance Map Data file could not be opened.\n"); return false; }
    action_count = (int)rows[0][0];
    for (size_t i = 1; i < rows.size(); ++i) if (!rows[i].empty()) data.push_back(rows[i]);
    return true;
}
bool load_vertices(const std::string& path, std::vector<float>& vertices) {
    std::vector<std::vector<float>> rows;
    if (!read_float_lines(path, rows)) { printf("Scene Data file could not be opened.\n"); return false; }
    for (const auto& r : rows) vertices.insert(vertices.end(), r.begin(), r.end());
    return true;
}
QValueTrainerResult train_q_value_network(Renderer& renderer, const std::string& data_path, const std::string& vertices_path, int epochs, int batch_size,
                                          const char* save_model, const char* load_model, unsigned shuffle_seed, bool verbose) {
    QValueTrainerResult res;
    std::vector<std::vector<float>> data; int action_count = 0;
    if (!load_radiance_map_data(data_path, data, action_count)) return res;
    // std::random_shuffle (main.cu:126) is gone from C++17; a Fisher-Yates shuffle on a seeded LCG takes its place (any permutation serves)
    { uint64_t st = 0x9E3779B97F4A7C15ull ^ shuffle_seed; for (size_t i = data.size(); i > 1; --i) { st = st * 6364136223846793005ull + 1442695040888963407ull; std::swap(data[i - 1], data[(size_t)((st >> 33) % i)]); } }
    std::vector<float> vertices;
    if (!load_vertices(vertices_path, vertices)) return res;
    res.lines = (int)data.size(); res.vertices = (int)vertices.size() / 3; res.action_count = action_count;
    if (verbose) { std::cout << "Read " << data.size() << " lines of radiance_map data." << std::endl << "Action count: " << action_count << std::endl << "Read " << vertices.size() / 3 << " vertices." << std::endl; }
    if (action_count != RLPT_GRID_CELLS || vertices.size() % 9 != 0 || vertices.empty()) throw RenderError{ RLPT_ERR_ARG, "train_q_value_network: expected 144 actions and 9 floats per triangle in vertices.txt" };
    std::vector<const std::vector<float>*> training, test;
    for (const auto& line : data) {
        if (line.size() != (size_t)(3 + action_count)) throw RenderError{ RLPT_ERR_IO, "train_q_value_network: a data line does not hold 3 + action_count values" };
        const double rv = (double)rand() / (RAND_MAX);                       // main.cu:147
        (rv < 0.8 ? training : test).push_back(&line);
    }
    res.train = (int)training.size(); res.test = (int)test.size();
    if (verbose) { std::cout << "Training data set size: " << training.size() << std::endl << "Test data set size: " << test.size() << std::endl; }
    // the network's input is every scene vertex minus the query point: the scene behind vertices.txt is uploaded as plain geometry
    rlpt_ctx* ctx = renderer.ctx();
    {
        const int n_tri = (int)vertices.size() / 9; std::vector<float> rgb(3 * (size_t)n_tri, 0.75f);
        renderer.check(rlpt_scene_upload(ctx, vertices.data(), rgb.data(), n_tri, nullptr, nullptr, 0));
        renderer.check(rlpt_dqn_set_vertices(ctx, vertices.data(), (int)vertices.size()));
    }
    if (load_model) renderer.check(rlpt_dqn_load_text(ctx, load_model)); else renderer.check(rlpt_dqn_init(ctx, 1984u));
    const size_t num_batches = (training.size() + (size_t)batch_size - 1) / (size_t)batch_size;
    std::vector<float> pos, tgt, q;
    for (int e = 0; e < epochs; ++e) {
        float loss = 0.f;
        for (size_t b = 0; b < num_batches; ++b) {
            const size_t sidx = b * (size_t)batch_size, n = std::min(training.size() - sidx, (size_t)batch_size);
            pos.resize(3 * n); tgt.resize((size_t)action_count * n);
            for (size_t k = 0; k < n; ++k) {
                const std::vector<float>& l = *training[sidx + k];
                std::copy(l.begin(), l.begin() + 3, pos.begin() + 3 * k); std::copy(l.begin() + 3, l.end(), tgt.begin() + (size_t)action_count * k);
            }
            float bl = 0.f;
            renderer.check(rlpt_dqn_train_supervised(ctx, pos.data(), tgt.data(), (int)n, 1, &bl));
            loss += bl;
        }
        float error = 0.f;
        if (!test.empty()) {
            pos.resize(3 * test.size()); q.resize((size_t)action_count * test.size());
            for (size_t t = 0; t < test.size(); ++t) std::copy(test[t]->begin(), test[t]->begin() + 3, pos.begin() + 3 * t);
            renderer.check(rlpt_dqn_forward(ctx, pos.data(), (int)test.size(), q.data()));
            for (size_t t = 0; t < test.size(); ++t) for (int a = 0; a < action_count; ++a) { const float d = (*test[t])[3 + a] - q[t * (size_t)action_count + a]; error += d * d; }
        }
        res.loss.push_back(loss); res.error.push_back(error);
        if (verbose) {
            std::cout << "---------------- " << e + 1 << " ----------------" << std::endl << "      Loss: " << loss << std::endl << "      Error: " << error << std::endl
                      << "-------------------------------------" << std::endl << std::endl;
        }
    }
    if (save_model) renderer.check(rlpt_dqn_save_text(ctx, save_model));
    return res;
}

}  // namespace rlpt_host

// ------------------------------------------------------------------------------------------------ C entry points for tests
// (ctypes cannot call C++): the loaders' output as flat arrays. Counts first (arrays may be NULL), then fill.
extern "C" int rlpt_host_load_scene(const char* obj_path_or_null, int lights_in_obj, int preset /* 0 committed, 1 door room, 2 normalised, no lights */,
                                    int* n_surfaces, int* n_lights, float* sv, float* srgb, float* snrm, float* slum, float* lv, float* lrgb, float* lnrm, float* llum,
                                    float* vertices, int* vertices_count) {
    using namespace rlpt_host;
    Scene s;
    if (!obj_path_or_null) s.load_cornell_box_scene();
    else {
        ImportPreset p = preset == 1 ? ImportPreset::door_room() : (preset == 2 ? ImportPreset::normalised_no_lights() : ImportPreset::committed());
        if (!s.load_custom_scene(obj_path_or_null, lights_in_obj != 0, p)) return 1;
    }
    if (n_surfaces) *n_surfaces = s.surfaces_count;
    if (n_lights) *n_lights = s.area_light_count;
    if (vertices_count) *vertices_count = s.vertices_count;
    auto put = [](float* v, float* n, int i, const Triangle& t) {
        if (v) { const float a[9] = { t.v0.x, t.v0.y, t.v0.z, t.v1.x, t.v1.y, t.v1.z, t.v2.x, t.v2.y, t.v2.z }; std::copy(a, a + 9, v + 9 * (size_t)i); }
        if (n) { n[3 * (size_t)i] = t.normal.x; n[3 * (size_t)i + 1] = t.normal.y; n[3 * (size_t)i + 2] = t.normal.z; }
    };
    for (int i = 0; i < s.surfaces_count; ++i) {
        put(sv, snrm, i, s.surfaces[i]);
        if (srgb) { srgb[3 * i] = s.surfaces[i].material.diffuse_c.x; srgb[3 * i + 1] = s.surfaces[i].material.diffuse_c.y; srgb[3 * i + 2] = s.surfaces[i].material.diffuse_c.z; }
        if (slum) slum[i] = s.surfaces[i].material.luminance;
    }
    for (int i = 0; i < s.area_light_count; ++i) {
        put(lv, lnrm, i, s.area_lights[i]);
        if (lrgb) { lrgb[3 * i] = s.area_lights[i].diffuse_p.x; lrgb[3 * i + 1] = s.area_lights[i].diffuse_p.y; lrgb[3 * i + 2] = s.area_lights[i].diffuse_p.z; }
        if (llum) llum[i] = s.area_lights[i].luminance;
    }
    if (vertices) std::copy(s.vertices, s.vertices + s.vertices_count, vertices);
    return 0;
}
extern "C" int rlpt_host_saved_volumes_to_surfaces(const char* path, int max_surfaces, float* sv, float* srgb, float* snrm) {
    using namespace rlpt_host;
    std::vector<Surface> S;
    if (!read_radiance_volumes_to_surfaces(path, S)) return -1;
    for (int i = 0; i < (int)S.size() && i < max_surfaces; ++i) {
        const Surface& s = S[i];
        const float v[9] = { s.v0.x, s.v0.y, s.v0.z, s.v1.x, s.v1.y, s.v1.z, s.v2.x, s.v2.y, s.v2.z };
        memcpy(sv + 9 * i, v, sizeof v);
        srgb[3 * i] = s.material.diffuse_c.x; srgb[3 * i + 1] = s.material.diffuse_c.y; srgb[3 * i + 2] = s.material.diffuse_c.z;
        snrm[3 * i] = s.normal.x; snrm[3 * i + 1] = s.normal.y; snrm[3 * i + 2] = s.normal.z;
    }
    return (int)S.size();
}
