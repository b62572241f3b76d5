"""Small runs of every tracing path (both tracers, BVH scene, Voronoi view, Neural-Q training and inference): a crash check"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200"))
import rlpt
z = np.load(os.path.join(ROOT, "tests/golden/scenes.npz"))
def scene(n): return {k.split("/")[1]: z[k] for k in z.files if k.startswith(n + "/")}
for name, cam in (("cornell", (0, 0, -3)), ("archway", (-1, 0.2, -0.99))):
    s = scene(name)
    c = rlpt.Context(0, width=48, height=40, spp=3, max_bounces=80)
    c.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); c.camera_set(cam)
    c.render_default(2)
    nv = c.radiance_map_build()
    c.render_sarsa(3); c.render_voronoi()
    img = c.frame_download(); st = c.stats()
    print(name, nv, float(img.mean()), st["paths"], st["kd_fallbacks"])
    if name == "cornell":
        c.dqn_init(3); c.configure(width=16, height=16, spp=1)
        print("nq loss", c.render_neuralq(1, batch=128)); c.render_pretrained(1)
    c.close()
print("done")
