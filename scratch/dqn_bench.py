import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200"))
import rlpt
z = np.load(os.path.join(ROOT, "tests/golden/scenes.npz")); name = sys.argv[1] if len(sys.argv) > 1 else "archway"
s = {k.split("/")[1]: z[k] for k in z.files if k.startswith(name + "/")}
c = rlpt.Context(0, width=512, height=512, spp=1, max_bounces=80); c.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); c.camera_set((-1, 0.2, -0.99) if name == "archway" else (0, 0, -3))
c.dqn_init(1984)
n = 262144
pos = np.random.RandomState(0).uniform(-1, 1, (n, 3)).astype(np.float32)
for rep in range(3):
    t = time.perf_counter(); q = c.dqn_forward(pos); dt = time.perf_counter() - t
print("forward incl. host copies: %.1f ms" % (dt * 1e3))
c.stats_reset(); t = time.perf_counter(); c.render_pretrained(1); c.sync(); dt = time.perf_counter() - t; st = c.stats()
print("pretrained frame (1 spp): %.1f ms, %.2f Mpaths/s, path length %.2f" % (dt * 1e3, st["paths"] / st["device_seconds"] / 1e6, st["path_length_sum"] / st["paths"]))
c.configure(spp=8); c.stats_reset(); c.render_pretrained(1); st = c.stats()
print("pretrained frame (8 spp): %.2f Mpaths/s" % (st["paths"] / st["device_seconds"] / 1e6))
c.configure(spp=1); c.stats_reset(); t = time.perf_counter(); loss = c.render_neuralq(1, batch=int(sys.argv[2]) if len(sys.argv) > 2 else 4096); dt = time.perf_counter() - t; st = c.stats()
print("neural-q training frame (1 spp, batch %s): %.1f ms wall, %.3f Mpaths/s, launches %d" % (sys.argv[2] if len(sys.argv) > 2 else "4096", dt * 1e3, st["paths"] / st["device_seconds"] / 1e6, st["kernel_launches"]))
