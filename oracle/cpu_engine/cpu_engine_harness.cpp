/* oracle/cpu_engine/cpu_engine_harness.cpp -- TEST / BASELINE INFRASTRUCTURE ONLY (never the product, never shipped).
 *
 * Driver + restoration patch for the reference's Old_CPU_Rendering_Engine (BASELINE.json configs[0]). The engine's
 * sources are compiled unmodified from /root/reference by oracle/cpu_engine/build.sh; this file adds only what is missing:
 *   1. the two functions the committed sources declare but never define (see cpu_engine_shim.h): Triangle::intersects and
 *      Triangle::cramer, restated from the GPU engine's equivalents G/rays/ray.cu:39-74 and :115-141 (same Cramer's-rule test,
 *      dir * SCREEN_HEIGHT scaling, t, u, v >= 0, u + v <= 1, strict < on the distance);
 *   2. a headless SDLScreen (C/sdl/sdl_screen.h: same members; PutPixelSDL keeps the colour conversion of C/sdl/sdl_screen.cpp:96-108);
 *   3. extern "C" entry points that do what C/main.cpp:63-135 does for PATH_TRACING_METHOD 3 and time draw_default_path_tracing
 *      (C/path_tracing/default_path_tracing.cpp:5-18) alone with std::chrono::steady_clock.
 */
#include <chrono>
#include <cstring>
#include <vector>
#include <omp.h>
#include <glm/glm.hpp>

#include "sdl_screen.h"
#include "image_settings.h"
#include "monte_carlo_settings.h"
#include "cornell_box_scene.h"
#include "ray.h"
#include "camera.h"
#include "area_light_plane.h"
#include "surface.h"
#include "triangle.h"
#include "default_path_tracing.h"

/* ---- 1. restoration: follows G/rays/ray.cu:115-141 */
bool Triangle::cramer(mat3 A, vec3 b, vec3& solution) {
    solution = vec3(0, 0, 0);
    const float detA = glm::determinant(A);
    if (detA == 0) return false;
    const mat3 keep = A;
    A[0] = b; solution.x = glm::determinant(A) / detA; A = keep;
    A[1] = b; solution.y = glm::determinant(A) / detA; A = keep;
    A[2] = b; solution.z = glm::determinant(A) / detA;
    return true;
}
/* follows G/rays/ray.cu:39-74 with the CPU types (C/rays/ray.h:38-44 Intersection, getters of C/rays/ray.h:69-70) */
bool Triangle::intersects(Ray* ray, Intersection& intersection, int index) {
    const vec4 start = ray->get_start();
    vec4 dir = ray->get_direction();
    const vec3 e1(v1.x - v0.x, v1.y - v0.y, v1.z - v0.z), e2(v2.x - v0.x, v2.y - v0.y, v2.z - v0.z), b(start.x - v0.x, start.y - v0.y, start.z - v0.z);
    dir = vec4(vec3(dir) * (float)SCREEN_HEIGHT, 1);
    const mat3 A(vec3(-dir), e1, e2);
    vec3 s;
    if (!cramer(A, b, s) || !(s.x >= 0.0f && s.y >= 0.0f && s.z >= 0.0f && s.y + s.z <= 1.0f)) return false;
    if (!(s.x < intersection.distance)) return false;
    intersection.position = start + s.x * dir; intersection.position[3] = 1;
    intersection.distance = s.x; intersection.normal = normal; intersection.index = index;
    return true;
}

/* ---- 2. headless SDLScreen */
SDLScreen::SDLScreen(int width, int height, bool) {
    this->width = width; this->height = height; this->window = nullptr; this->renderer = nullptr; this->texture = nullptr;
    this->buffer = new uint32_t[(size_t)width * height];
    memset(this->buffer, 0, (size_t)width * height * sizeof(uint32_t));
}
void SDLScreen::kill_screen() { delete[] this->buffer; this->buffer = nullptr; }
void SDLScreen::SDL_Renderframe() {}
void SDLScreen::SDL_SaveImage(const char*) {}
bool SDLScreen::NoQuitMessageSDL() { return true; }
void SDLScreen::PutPixelSDL(int x, int y, glm::vec3 colour) {
    if (x < 0 || x >= this->width || y < 0 || y >= this->height) return;
    const uint32_t r = uint32_t(glm::clamp(255 * colour.r, 0.f, 255.f)), g = uint32_t(glm::clamp(255 * colour.g, 0.f, 255.f)), b = uint32_t(glm::clamp(255 * colour.b, 0.f, 255.f));
    this->buffer[y * this->width + x] = (128u << 24) + (r << 16) + (g << 8) + b;
}

/* ---- 3. driver */
extern "C" {
int cpu_engine_dims(int* w, int* h, int* spp, int* bounces) { *w = SCREEN_WIDTH; *h = SCREEN_HEIGHT; *spp = SAMPLES_PER_PIXEL; *bounces = MAX_RAY_BOUNCES; return 0; }
int cpu_engine_max_threads(void) { return omp_get_num_procs(); }
/* renders `frames` frames of the built-in Cornell box with `threads` OpenMP threads (the reference hard-codes 6, C/main.cpp:65);
 * seconds[f] = time inside draw_default_path_tracing for frame f; argb (may be null) receives the last frame, W*H words */
int cpu_engine_render_default(int frames, int threads, double* seconds, uint32_t* argb) {
    omp_set_num_threads(threads > 0 ? threads : 6);
    SDLScreen screen(SCREEN_WIDTH, SCREEN_HEIGHT, FULLSCREEN_MODE);
    std::vector<Surface> surfaces_load; std::vector<AreaLightPlane> light_planes_load;
    get_cornell_shapes(surfaces_load, light_planes_load);
    Camera camera = Camera(vec4(0, 0, -3, 1));                                   /* C/main.cpp:77 */
    std::vector<Surface*> surfaces; for (size_t i = 0; i < surfaces_load.size(); i++) surfaces.push_back(&surfaces_load[i]);
    std::vector<AreaLightPlane*> light_planes; for (size_t i = 0; i < light_planes_load.size(); i++) light_planes.push_back(&light_planes_load[i]);
    for (int f = 0; f < frames; ++f) {
        const auto t0 = std::chrono::steady_clock::now();
        draw_default_path_tracing(screen, camera, light_planes, surfaces);
        seconds[f] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (argb) memcpy(argb, screen.buffer, sizeof(uint32_t) * (size_t)SCREEN_WIDTH * SCREEN_HEIGHT);
    screen.kill_screen();
    return 0;
}
}
