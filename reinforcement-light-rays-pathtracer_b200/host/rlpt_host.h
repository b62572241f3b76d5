// rlpt_host.h -- C++ host mirror of the reference engine's surface for the hot path, written purely against the
// C ABI in include/rlpt.h. A program shaped like the reference's main.cu (G/main.cu:91-516) keeps its structure:
//   Scene scene; scene.load_cornell_box_scene() | scene.load_custom_scene(path, lights_in_obj); scene.save_vertices_to_file();
//   Camera camera(vec4(0, 0, -3, 1));  SDLScreen screen(W, H);
//   Renderer r(device); r.upload(scene); r.set_camera(camera);
//   RadianceMap radiance_map(r);                      // builds volumes + kd-tree + device tables (G/main.cu:264-289)
//   r.render_sarsa(1); r.present(screen); screen.SDL_SaveImage("render.bmp");
// Same names, argument meaning and file formats as the reference classes they mirror (file:line at each declaration);
// what the reference fixes with #defines (G/constants/) is a run-time Settings value here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rlpt.h"

namespace rlpt_host {

struct vec3 { float x = 0, y = 0, z = 0; vec3() {} vec3(float a) : x(a), y(a), z(a) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct vec4 { float x = 0, y = 0, z = 0, w = 0; vec4() {} vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {} };

// G/objects/material.cuh:17-28
struct Material { vec3 diffuse_c; float luminance = 0; Material() {} explicit Material(vec3 c); };
// G/objects/triangle.cuh:13-42
struct Triangle { vec4 v0, v1, v2, normal; Triangle() {} Triangle(vec4 a, vec4 b, vec4 c) : v0(a), v1(b), v2(c) {} void compute_and_set_normal(); float compute_area() const; };
// G/objects/surface.cuh:18-32
struct Surface : Triangle { Material material; Surface() {} Surface(vec4 a, vec4 b, vec4 c, Material m) : Triangle(a, b, c), material(m) {} };
// G/lights/area_light.cuh:21-36
struct AreaLight : Triangle { vec3 diffuse_p; float luminance = 0; AreaLight() {} AreaLight(vec4 a, vec4 b, vec4 c, vec3 p); };

// What G/objects/object_importer.cu hard-codes per scene (materials by triangle index :150-163, lights :212-271, the
// lights-in-obj index ranges :375-389, the commented normalisation :119). `committed()` is the file as committed.
struct ImportPreset {
    enum Lights { ARCHWAY, DOOR_ROOM, SIMPLE_CLOSED_ROOM, SIMPLE_ROOM, NONE } lights = ARCHWAY;
    enum Colours { COMMITTED /* red i>80, blue 11<i<24 */, DOOR_ROOM_COLOURS /* red 23<i<36, blue 11<i<24 */, ALL_WHITE } colours = COMMITTED;
    bool normalise = false;                 // scale = 2 / max_difference instead of 2 (needed for Medieval_House)
    static ImportPreset committed() { return ImportPreset(); }
    static ImportPreset door_room() { ImportPreset p; p.lights = DOOR_ROOM; p.colours = DOOR_ROOM_COLOURS; return p; }
    static ImportPreset normalised_no_lights() { ImportPreset p; p.lights = NONE; p.colours = ALL_WHITE; p.normalise = true; return p; }
};

// G/scenes/scene.cuh:27-47
class Scene {
public:
    Surface* surfaces = nullptr; int surfaces_count = 0;
    AreaLight* area_lights = nullptr; int area_light_count = 0;
    float* vertices = nullptr; int vertices_count = 0;          // 9 floats per triangle, surfaces then lights
    Scene() {}
    ~Scene();
    Scene(const Scene&) = delete; Scene& operator=(const Scene&) = delete;
    void load_cornell_box_scene();                                               // G/scenes/scene.cu:8-30
    bool load_custom_scene(const char* filename, bool lights_in_obj, const ImportPreset& preset = ImportPreset::committed(),
                           const char* saved_radiance_volumes = nullptr /* RENDER_SAVED_RADIANCE_VOLUMES: file whose volumes are appended as geometry, :41-46 */);   // :33-60
    void save_vertices_to_file(const char* path = "../Radiance_Map_Data/vertices.txt") const;   // :63-88
private:
    void adopt(std::vector<Surface>& s, std::vector<AreaLight>& l, std::vector<float>& v);
};
// G/objects/object_importer.cu:8-89; returns false when the file cannot be opened (the reference prints and carries on)
bool load_scene(const char* path, std::vector<Surface>& surfaces, std::vector<AreaLight>& area_lights, std::vector<float>& vertices,
                bool lights_in_obj, const ImportPreset& preset = ImportPreset::committed());
void get_cornell_shapes(std::vector<Surface>& surfaces, std::vector<AreaLight>& lights, std::vector<float>& vertices);   // G/scenes/cornell_box_scene.cu:4-245

// G/camera.cuh:11-36 (yaw_x is initialised here; the reference leaves it indeterminate, G/camera.cu:3-7)
class Camera {
public:
    vec4 position; float yaw_y = 0.f, yaw_x = 0.f; float R[4][4];
    explicit Camera(vec4 position);
    void rotate_left(float y); void rotate_right(float y); void rotate_up(float x); void rotate_down(float x);
    void move_forwards(float distance); void move_backwards(float distance);
};

// G/sdl/sdl_screen.h: headless equivalent (no window): PutPixelSDL colour conversion and the BMP SDL_SaveBMP writes
class SDLScreen {
public:
    int width, height; std::vector<uint32_t> buffer;
    SDLScreen(int width, int height) : width(width), height(height), buffer((size_t)width * height, 0u) {}
    void PutPixelSDL(int x, int y, vec3 colour);                // G/sdl/sdl_screen.cpp:96-108
    void SDL_Renderframe() {}                                   // :68-76 (nothing to present without a window)
    bool SDL_SaveImage(const char* filename) const;             // :60-66
};

struct RenderError { int status; std::string message; };

// The frame loops of G/main.cu (method 0 :207-244, method 1 :301-364) over one rlpt_ctx. Throws RenderError.
class Renderer {
public:
    explicit Renderer(int device = 0);
    ~Renderer();
    Renderer(const Renderer&) = delete; Renderer& operator=(const Renderer&) = delete;
    rlpt_ctx* ctx() const { return ctx_; }
    rlpt_config& settings() { return cfg_; }                    // edit, then apply_settings()
    void apply_settings();
    void upload(const Scene& scene);                            // G/main.cu:161-186
    void set_camera(const Camera& camera);                      // G/main.cu:210,307
    void render_default(int frames = 1);                        // draw_default_path_tracing<<<>>> per frame
    void render_sarsa(int frames = 1);                          // draw_reinforcement_path_tracing<<<>>> + update_radiance_volume_distributions<<<>>>
    void reset_frame();                                         // cudaMemset(device_buffer, 0) (G/main.cu:241,359)
    void download(std::vector<float>& rgb);                     // cudaMemcpy(host_buffer, device_buffer) (G/main.cu:232,349)
    void present(SDLScreen& screen);                            // the PutPixelSDL loop (G/main.cu:235-239)
    rlpt_stats_t stats();
    void append_training_stats(const char* path);               // "avg_path_length 0.0 zero_contribution_paths" (G/main.cu:336-339)
    void check(int status) const;
private:
    rlpt_ctx* ctx_ = nullptr; rlpt_config cfg_{};
};

// G/radiance_volumes/radiance_map.cuh:30-83: lifecycle + files. Construction builds the volumes, the kd-tree and the
// device tables (rlpt_radiance_map_build); the sampling/TD methods are device code inside the library.
class RadianceMap {
public:
    int radiance_volumes_count = 0, radiance_array_size = 0;
    explicit RadianceMap(Renderer& r);
    void update_radiance_volume_distributions();                               // reinforcement_path_tracing.cu:6-13
    void save_q_vals_to_file(const char* path = "../Radiance_Map_Data/radiance_map_data.txt");        // radiance_map.cu:237-268
    void load_q_vals_from_file(const char* path);                              // loader the reference lacks
    void save_selected_radiance_volumes_vals(std::string fpath);               // radiance_map.cu:272-302 (to_select.txt -> selected_sarsa.txt)
private:
    Renderer& r_;
};
// G/deep_learning/neural_q_pathtracer.cuh:88-96. The reference's constructor does everything (initialise DyNet, optionally
// load the model, render `frames` frames with training, save the model after every frame, save the image). Same here, over
// rlpt_render_neuralq; argc/argv (DyNet flags) are accepted and ignored.
class NeuralQPathtracer {
public:
    NeuralQPathtracer(unsigned int frames, int batch_size, SDLScreen& screen, Renderer& renderer, Scene& scene, Camera& camera, int argc = 0, char** argv = nullptr,
                      const char* load_model = nullptr /* LOAD_MODEL */, const char* save_model = "../Radiance_Map_Data/deep_q_learning_12_12.model" /* SAVE_MODEL */,
                      const char* stats_file = "../Radiance_Map_Data/nn_training_stats.txt", const char* image = "../Images/render.bmp");
    double last_loss = 0.0;
};
// G/deep_learning/pre_trained_pathtracer.cuh:92-100: render with a trained network, no learning; returns silently when the
// model file is missing, as the reference does (pre_trained_pathtracer.cu:50-53)
class PretrainedPathtracer {
public:
    PretrainedPathtracer(unsigned int frames, int batch_size, SDLScreen& screen, Renderer& renderer, Scene& scene, Camera& camera, int argc = 0, char** argv = nullptr,
                         const char* model = "../Radiance_Map_Data/deep_q_learning_12_12.model", const char* image = "../Images/render.bmp");
    bool rendered = false;
};
// NN_Q_Value_Trainer/Source/main.cu:118-295 -- the offline supervised trainer: fit the network to a saved Q table. Reads the two files
// the engine writes (radiance_map_data.txt: "144" then "px py pz q0 .. q143" per volume, RadianceMap::save_q_vals_to_file; vertices.txt,
// Scene::save_vertices_to_file), shuffles, splits ~80 / 20 with rand() (:143-155), trains EPOCHS x ceil(train / BATCH_SIZE) Adam steps on the
// summed squared distance over all outputs (:186-238, rlpt_dqn_train_supervised), evaluates the test error after every epoch (:240-279),
// prints the reference's per-epoch block and saves the DyNet text model (:287-292). settings.cuh: BATCH_SIZE 128, EPOCHS 100.
struct QValueTrainerResult { int lines = 0, vertices = 0, train = 0, test = 0, action_count = 0; std::vector<float> loss, error; };
bool load_radiance_map_data(const std::string& path, std::vector<std::vector<float>>& radiance_map_data, int& action_count);      // main.cu:72-116
bool load_vertices(const std::string& path, std::vector<float>& vertices);                                                          // main.cu:39-69
QValueTrainerResult train_q_value_network(Renderer& renderer, const std::string& radiance_map_data = "../Radiance_Map_Data/radiance_map_data.txt",
                                          const std::string& vertices = "../Radiance_Map_Data/vertices.txt", int epochs = 100, int batch_size = 128,
                                          const char* save_model = "../Radiance_Map_Data/radiance_map_model.model", const char* load_model = nullptr,
                                          unsigned shuffle_seed = 0, bool verbose = true);
// Saved radiance volumes drawn as geometry (RENDER_SAVED_RADIANCE_VOLUMES, G/constants/image_settings.h:15, G/scenes/scene.cu:41-46):
// each volume of a selected_*.txt file becomes a hemisphere of 12x12 quads (two Surfaces each) of diameter DIAMETER, coloured
// from green to red by its distribution value relative to the volume's maximum.
struct SavedRadianceVolume {
    vec4 position; vec3 normal; float radiance_distribution[RLPT_GRID_CELLS];
    std::vector<std::vector<vec4>> get_vertices() const;                 // G/radiance_volumes/radiance_volume.cu:441-463
    void build_surfaces(std::vector<Surface>& surfaces) const;           // :467-496
};
bool read_radiance_volumes_from_file(const std::string& fname, std::vector<SavedRadianceVolume>& rvs);       // :377-437 ("px py pz nx ny nz d0 .. d143" per line)
bool read_radiance_volumes_to_surfaces(const std::string& fname, std::vector<Surface>& surfaces);            // :499-515
// Shirley-Chiu square -> unit hemisphere (y up), G/utils/hemisphere_helpers.cu:134-226
void map(float x, float y, float& x_ret, float& y_ret, float& z_ret);
// G/utils/hemisphere_helpers.cu:230-281: "px py pz nx ny nz" per line
bool read_hemisphere_locations_and_normals(const std::string& path, std::vector<vec3>& locations, std::vector<vec3>& normals);

}  // namespace rlpt_host

// ---- NCCL all-reduce hook for rlpt_set_allreduce (host/nccl_hook.cpp); `comm` is an ncclComm_t
extern "C" int rlpt_nccl_allreduce(void* d_buf, uint64_t count, int dtype, void* cuda_stream, void* comm);
