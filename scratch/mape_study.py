import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rlpt
from checkers import Reference, mape_score, to_rgb8
z = np.load(os.path.join(ROOT, "tests/golden/scenes.npz")); s = {k.split("/")[1]: z[k] for k in z.files if k.startswith("cornell/")}
R = Reference("cuda"); w, h, spp = R.width, R.height, R.spp
img8 = lambda rgb: to_rgb8(np.asarray(rgb, np.float32).reshape(w, h, 3).transpose(1, 0, 2))
R.scene_arrays(s["sv"], s["srgb"], s["lv"], s["lrgb"]); R.camera(0, 0, -3)
gtA, _ = R.render_default(32); gtB, _ = R.render_default(32)
ref32 = [R.render_default(1)[0] for _ in range(6)]
c = rlpt.Context(0, width=w, height=h, spp=spp, max_bounces=80); c.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); c.camera_set((0, 0, -3))
prod32 = []
for i in range(6):
    c.frame_reset(); c.render_default(1); prod32.append(c.frame_download().copy())
c.frame_reset(); c.render_default(32); p1024 = c.frame_download().copy()
out = dict(ref32=[mape_score(img8(gtA), img8(x)) for x in ref32], prod32=[mape_score(img8(gtA), img8(x)) for x in prod32],
           refB_vs_refA=mape_score(img8(gtA), img8(gtB)), prod1024_vs_refA=mape_score(img8(gtA), img8(p1024)), prod1024_vs_refB=mape_score(img8(gtB), img8(p1024)),
           mean_refA=np.nan_to_num(gtA).mean(0).tolist(), mean_refB=np.nan_to_num(gtB).mean(0).tolist(), mean_prod=p1024.mean(0).tolist(),
           nan_refA=int(np.isnan(gtA).any(1).sum()))
# relative error on bright pixels only
gA, gB, pp = np.nan_to_num(gtA), np.nan_to_num(gtB), p1024
m = gA.max(1) > 0.05
out["rel_l1_bright_refB"] = float(np.abs(gB[m] - gA[m]).mean() / gA[m].mean()); out["rel_l1_bright_prod"] = float(np.abs(pp[m] - gA[m]).mean() / gA[m].mean())
# 16x16 block means
blk = lambda a: a.reshape(w // 16, 16, h // 16, 16, 3).mean((1, 3))
out["blk_rel_refB"] = float(np.abs(blk(gB) - blk(gA)).mean() / blk(gA).mean()); out["blk_rel_prod"] = float(np.abs(blk(pp) - blk(gA)).mean() / blk(gA).mean())
print(json.dumps(out, indent=1))
