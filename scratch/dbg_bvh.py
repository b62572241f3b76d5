import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rlpt
from checkers import Oracle, build_oracle
build_oracle()
z = np.load(os.path.join(ROOT, "tests/golden/scenes.npz")); s = {k.split("/")[1]: z[k] for k in z.files if k.startswith("archway/")}
w = h = 64; spp = 8; cam = (-1.0, 0.2, -0.99)
for trav in (0, 1, 2):
    c = rlpt.Context(0, width=w, height=h, spp=spp, max_bounces=80, traversal=trav); c.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); c.camera_set(cam)
    o = Oracle(); o.scene_set(s["sv"], s["srgb"], s["lv"], s["lrgb"])
    c.render_default(1); img = c.frame_download()
    oi, st = o.render_frame(0, w, h, spp, sample0=0, max_bounces=80, cam=cam, fma_mode=1, td_mode=1); oi = (oi / spp).astype(np.float32)
    err = np.abs(img - oi).max(1) / np.maximum(np.abs(oi).max(1), 1e-2)
    bad = np.nonzero(err > 2e-3)[0]
    print("traversal", trav, "bad", len(bad), "of", w * h, "stats", c.stats()["path_length_sum"], st["total_path_length"], "mean", img.mean(), oi.mean())
    for b in bad[:8]:
        print("  pixel", b // h, b % h, img[b], oi[b])
    c.close()
