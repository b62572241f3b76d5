// example_main.cpp -- the reference's main.cu flow (G/main.cu:91-516) on the host mirror: load a scene, train + render
// with Expected-SARSA radiance volumes (or the default path tracer), save the BMP and the training statistics.
// With more than one GPU (--gpus N) one thread drives each GPU and the Q accumulators are all-reduced over NCCL.
//   rlpt_example [--scene cornell|PATH.obj] [--lights-in-obj] [--preset committed|door_room|normalised] [--method 0|1|3|4]
//                [--frames F] [--spp S] [--size W] [--camera x y z] [--env E] [--gpus N] [--out render.bmp]
//                [--save-q radiance_map_data.txt] [--save-vertices vertices.txt]       (method 1: RadianceMap::save_q_vals_to_file, Scene::save_vertices_to_file)
//                [--batch B] [--model FILE.model]                                      (method 3 NeuralQPathtracer: saves FILE; method 4 PretrainedPathtracer: loads it)
//   rlpt_example --train-q radiance_map_data.txt vertices.txt OUT.model [--epochs E] [--batch B]      (NN_Q_Value_Trainer/Source/main.cu)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <nccl.h>
#include "rlpt_host.h"

using namespace rlpt_host;

int main(int argc, char** argv) {
    std::string scene_arg = "cornell", preset_arg = "committed", out = "render.bmp";
    bool lights_in_obj = false; int method = 1, frames = 8, spp = 32, size = 512, gpus = 1; float cam[3] = { 0.f, 0.f, -3.f }, env = 0.f;
    std::string save_q, save_vertices, model, train_data, train_vertices, train_out; int batch = 0, epochs = 100;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : (char*)""; };
        if (a == "--scene") scene_arg = next(); else if (a == "--lights-in-obj") lights_in_obj = true; else if (a == "--preset") preset_arg = next();
        else if (a == "--method") method = atoi(next()); else if (a == "--frames") frames = atoi(next()); else if (a == "--spp") spp = atoi(next());
        else if (a == "--size") size = atoi(next()); else if (a == "--env") env = (float)atof(next()); else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--save-q") save_q = next(); else if (a == "--save-vertices") save_vertices = next(); else if (a == "--model") model = next();
        else if (a == "--batch") batch = atoi(next()); else if (a == "--epochs") epochs = atoi(next());
        else if (a == "--train-q") { train_data = next(); train_vertices = next(); train_out = next(); }
        else if (a == "--out") out = next(); else if (a == "--camera") { for (int k = 0; k < 3; ++k) cam[k] = (float)atof(next()); }
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    try {
        if (!train_data.empty()) {                                 // the offline supervised trainer (NN_Q_Value_Trainer)
            Renderer r(0);
            QValueTrainerResult res = train_q_value_network(r, train_data, train_vertices, epochs, batch > 0 ? batch : 128, train_out.c_str());
            if (res.loss.empty()) return 1;
            printf("trained %d epochs on %d lines (%d test): loss %g -> %g, test error %g -> %g; wrote %s\n", (int)res.loss.size(), res.train, res.test, res.loss.front(), res.loss.back(),
                   res.error.front(), res.error.back(), train_out.c_str());
            return 0;
        }
        Scene scene;
        if (scene_arg == "cornell") scene.load_cornell_box_scene();
        else {
            ImportPreset p = preset_arg == "door_room" ? ImportPreset::door_room() : (preset_arg == "normalised" ? ImportPreset::normalised_no_lights() : ImportPreset::committed());
            if (!scene.load_custom_scene(scene_arg.c_str(), lights_in_obj, p)) return 1;
        }
        Camera camera(vec4(cam[0], cam[1], cam[2], 1.f));
        SDLScreen screen(size, size);
        if (!save_vertices.empty()) scene.save_vertices_to_file(save_vertices.c_str());
        if (method == 3 || method == 4) {                          // the Neural-Q tracers: their constructors do everything, as in the reference
            Renderer r(0);
            rlpt_config& s = r.settings(); s.width = size; s.height = size; s.spp = spp; s.env_light = env; r.apply_settings();
            r.upload(scene); r.set_camera(camera);
            if (method == 3) {
                NeuralQPathtracer t((unsigned)frames, batch > 0 ? batch : 4096, screen, r, scene, camera, 0, nullptr, nullptr, model.empty() ? nullptr : model.c_str(), nullptr, out.c_str());
                printf("Neural-Q: %d training frames, last loss %g\n", frames, t.last_loss);
            } else {
                PretrainedPathtracer t((unsigned)frames, batch > 0 ? batch : 4096, screen, r, scene, camera, 0, nullptr, model.c_str(), out.c_str());
                if (!t.rendered) { fprintf(stderr, "model file %s missing\n", model.c_str()); return 1; }
            }
            printf("wrote %s\n", out.c_str());
            return 0;
        }
        std::vector<ncclComm_t> comms(gpus);
        if (gpus > 1) { std::vector<int> devs(gpus); for (int g = 0; g < gpus; ++g) devs[g] = g; if (ncclCommInitAll(comms.data(), gpus, devs.data()) != ncclSuccess) { fprintf(stderr, "ncclCommInitAll failed\n"); return 1; } }
        std::vector<std::string> errors(gpus);
        auto worker = [&](int g) {
            try {
                Renderer r(g);
                rlpt_config& s = r.settings(); s.width = size; s.height = size; s.spp = spp; s.env_light = env; s.rank = g; s.world_size = gpus; r.apply_settings();
                r.upload(scene); r.set_camera(camera);
                if (gpus > 1) r.check(rlpt_set_allreduce(r.ctx(), rlpt_nccl_allreduce, (void*)comms[g]));
                if (method == 1) {
                    RadianceMap radiance_map(r);
                    if (g == 0) printf("%d radiance volumes, kd-tree of %d elements\n", radiance_map.radiance_volumes_count, radiance_map.radiance_array_size);
                    for (int f = 0; f < frames; ++f) { r.render_sarsa(1); if (g == 0) r.append_training_stats("sarsa_training_stats.txt"); }
                    if (g == 0 && !save_q.empty()) radiance_map.save_q_vals_to_file(save_q.c_str());
                } else r.render_default(frames);
                r.check(rlpt_frame_allreduce(r.ctx()));
                if (g == 0) {
                    r.present(screen);
                    rlpt_stats_t st = r.stats();
                    printf("%.0f paths, mean path length %.3f, %.3f s on the device, %.1f Mpaths/s per GPU\n", st.paths, st.path_length_sum / st.paths, st.device_seconds, st.paths / st.device_seconds / 1e6);
                }
            } catch (const RenderError& e) { errors[g] = e.message; }
        };
        std::vector<std::thread> th;
        for (int g = 0; g < gpus; ++g) th.emplace_back(worker, g);
        for (auto& t : th) t.join();
        for (int g = 0; g < gpus; ++g) if (!errors[g].empty()) { fprintf(stderr, "GPU %d: %s\n", g, errors[g].c_str()); return 1; }
        if (!screen.SDL_SaveImage(out.c_str())) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
        printf("wrote %s\n", out.c_str());
    } catch (const RenderError& e) { fprintf(stderr, "rlpt error %d: %s\n", e.status, e.message.c_str()); return 1; }
    return 0;
}
