#!/usr/bin/env python
"""scratch/mape_ablation.py -- attributes the Expected-SARSA MAPE difference between the product and the reference's own kernels
(VERDICT r1, weak item 2). Runs on a GPU box: ground truth = the reference's default tracer (its CUDA kernels) at 1024 spp; then the
8 x 32 spp SARSA protocol of tests/test_gpu_mape.py for
  * the reference's kernels (3 runs: its TD update is a racy read-modify-write, so it varies run to run),
  * the product,
  * the CPU oracle in four configurations that isolate the product's two deliberate deviations:
      td_mode 1 / 0          batched (Jacobi) TD merge at frame end   vs   the reference's in-place running mean (serialised here)
      inclusive / exclusive  proper initial CDF + clamped last bin    vs   the reference's initial k/144 CDF, failing samples = NaN
    all four trace the same Philox paths as the product, so oracle(td 1, inclusive) must reproduce the product.
Writes profiles/r2_sarsa_mape_ablation.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rlpt
from checkers import Oracle, Reference, mape_score, to_rgb8
z = np.load(os.path.join(ROOT, "tests/golden/scenes.npz")); s = {k.split("/")[1]: z[k] for k in z.files if k.startswith("cornell/")}
R = Reference("cuda"); w, h, spp = R.width, R.height, R.spp
img8 = lambda rgb: to_rgb8(np.asarray(rgb, np.float32).reshape(w, h, 3).transpose(1, 0, 2))
R.scene_arrays(s["sv"], s["srgb"], s["lv"], s["lrgb"]); R.camera(0, 0, -3)
gt = img8(R.render_default(32)[0])
FR = int(os.environ.get("ABL_FRAMES", "8"))
out = {"protocol": "Cornell 512x512, frame 0 dropped, mean of %d frames x %d spp, MAPE (Graphing/mape.py) vs the reference's default tracer at 1024 spp" % (FR, spp)}
ref_runs = []
for _ in range(3):
    R.rmap_build(); rmean, _, rst = R.render_sarsa(FR + 1, 1); ref_runs.append(mape_score(gt, img8(rmean)))
out["reference_kernels"] = {"runs": ref_runs, "median": float(np.median(ref_runs)), "nan_pixels_frame0": float(rst[0, 4]), "nan_pixels_last": float(rst[-1, 4])}
c = rlpt.Context(0, width=w, height=h, spp=spp, max_bounces=80); c.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); c.camera_set((0, 0, -3))
c.radiance_map_build(); c.render_sarsa(1); c.frame_reset(); c.render_sarsa(FR)
out["product"] = mape_score(gt, img8(c.frame_download())); c.close()
orc = Oracle(); orc.scene_set(s["sv"], s["srgb"], s["lv"], s["lrgb"])
for td_mode in (1, 0):
    for inclusive in (1, 0):
        t0 = time.time()
        orc.rmap_build()
        if inclusive:
            orc.rmap_update_distributions(); orc.rmap_merge_frame()
        acc = np.zeros((w * h, 3), np.float64); cnt = np.zeros(w * h, np.int64); failed = 0.0
        for f in range(FR + 1):
            o, st = orc.render_frame(1, w, h, spp, sample0=f * spp, max_bounces=80, cam=(0, 0, -3), fma_mode=1, td_mode=td_mode, clamp_last_bin=inclusive)
            if td_mode == 1:
                orc.rmap_merge_frame()
            orc.rmap_update_distributions()
            failed += st["failed"] if f else 0.0
            if f == 0:
                continue
            o = o / spp; ok = np.isfinite(o).all(1)
            acc[ok] += o[ok]; cnt[ok] += 1
        img = (acc / np.maximum(cnt, 1)[:, None]).astype(np.float32)
        key = "oracle_td%d_%s" % (td_mode, "inclusive_cdf" if inclusive else "exclusive_cdf_as_reference")
        out[key] = {"mape": mape_score(gt, img8(img)), "failed_samples_after_frame0": failed, "seconds": time.time() - t0}
        print(key, out[key], flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_sarsa_mape_ablation.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
