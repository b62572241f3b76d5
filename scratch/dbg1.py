import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rlpt
from checkers import Oracle
z = np.load(os.path.join(ROOT, "tests/golden/scenes.npz")); s = {k.split("/")[1]: z[k] for k in z.files if k.startswith("cornell/")}
def mk(w, h, spp):
    c = rlpt.Context(0, width=w, height=h, spp=spp, max_bounces=80); c.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); c.camera_set((0, 0, -3)); c.radiance_map_build(); return c
# 3. e2e breakdown
import torch
c = mk(512, 512, 32); c.render_sarsa(3)
pinned = torch.empty((512 * 512, 3), dtype=torch.float32, pin_memory=True).numpy()
T = {"cam": 0, "render": 0, "dl": 0, "stats": 0}
for i in range(8):
    t = time.perf_counter(); c.camera_set((0, 0, -3)); T["cam"] += time.perf_counter() - t
    t = time.perf_counter(); c.render_sarsa(1); T["render"] += time.perf_counter() - t
    t = time.perf_counter(); c.frame_download(pinned); T["dl"] += time.perf_counter() - t
    t = time.perf_counter(); st = c.stats(); T["stats"] += time.perf_counter() - t
print({k: v / 8 * 1e3 for k, v in T.items()}, "ms per step; device s/frame", st["device_seconds"] / st["frames"])
t = time.perf_counter(); c.render_sarsa(8); print("render 8 back to back ms/frame", (time.perf_counter() - t) / 8 * 1e3)
c.close()
