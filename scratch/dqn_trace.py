import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200"))
import rlpt
z = np.load(os.path.join(ROOT, "tests/golden/scenes.npz")); name = "cornell"
s = {k.split("/")[1]: z[k] for k in z.files if k.startswith(name + "/")}
c = rlpt.Context(0, width=512, height=512, spp=1, max_bounces=80); c.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); c.camera_set((0, 0, -3))
c.dqn_init(1984)
n = 262144
pos = np.random.RandomState(0).uniform(-1, 1, (n, 3)).astype(np.float32)
q = c.dqn_forward(pos); c.sync()
