"""-m gpu: image-level parity with the REFERENCE's own kernels (oracle/_ref/libref_cuda.so = G/path_tracing/*.cu etc.
compiled unmodified for sm_100a), judged with the reference's own metric, MAPE (Graphing/mape.py:10-21) on 8-bit RGB
after the PutPixelSDL conversion. Protocol: ground truth = the reference's default path tracer at
1024 spp; at matched spp the product's MAPE against that ground truth must equal the reference's own MAPE within the
stated tolerance, for the default tracer and for Expected SARSA. Random streams differ (XORWOW vs Philox), so this is
statistical by construction. Also records the reference kernels' own throughput on this GPU (printed, -s to see)."""
import json
import os

import numpy as np
import pytest

from conftest import CONFIG_SCENES, ROOT, load_scene

pytestmark = pytest.mark.gpu
TOL_DMAPE_32 = 0.04         # |mean MAPE(product) - mean MAPE(reference)| over 6 independent 32-spp frames (single-frame sigma ~0.03, measured)
TOL_DMAPE_1024 = 0.02       # |MAPE(product 1024 spp vs ref A) - MAPE(ref B vs ref A)|, A and B independent reference renders at 1024 spp
TOL_MEAN = 5e-3             # relative difference of the per-channel image means at 1024 spp
# Expected SARSA, 8 x 32 spp after one dropped frame. The product's two deliberate deviations both LOWER its error, and by how much is measured
# (profiles/r2_sarsa_mape_ablation.json, scratch/mape_ablation.py: the CPU oracle tracing the product's own Philox paths in the four combinations):
# batched TD merge vs the in-place update -0.029..-0.035, proper initial CDF + clamped last bin vs the reference's k/144 start -0.008..-0.015;
# oracle(batched, proper CDF) = product to 3e-4, and the reference's kernels (racy in-place update) land between oracle(batched, k/144) and
# oracle(in-place, k/144). So product - reference must lie in [-0.045, +0.012]: the sum of both effects on one side, the reference's own
# run-to-run spread (0.5213 .. 0.5310 over six runs) on the other.
TOL_DMAPE_SARSA_BETTER, TOL_DMAPE_SARSA_WORSE = 0.045, 0.012


def _img8(rgb, w, h):
    from checkers import to_rgb8
    return to_rgb8(np.asarray(rgb, np.float32).reshape(w, h, 3).transpose(1, 0, 2))


def test_mape_parity_cornell(ctx, ref_cuda, golden_scenes):
    from checkers import mape_score
    w, h, spp = ref_cuda.width, ref_cuda.height, ref_cuda.spp
    s = golden_scenes["cornell"]
    ref_cuda.scene_arrays(s["sv"], s["srgb"], s["lv"], s["lrgb"]); ref_cuda.camera(0.0, 0.0, -3.0)
    gt_f, st = ref_cuda.render_default(32)                                   # ground truth A: reference method 0 at 1024 spp
    gt = _img8(gt_f, w, h)
    gtb_f, _ = ref_cuda.render_default(32)                                   # independent reference render B (its RNG stream continues)
    mape_refb = mape_score(gt, _img8(gtb_f, w, h))
    ref32 = [ref_cuda.render_default(1) for _ in range(6)]
    mape_ref_default = float(np.mean([mape_score(gt, _img8(x, w, h)) for x, _ in ref32]))
    ref_default_mpaths = w * h * spp / (np.mean([t[0, 1] for _, t in ref32]) * 1e-3) / 1e6

    load_scene(ctx, s)
    ctx.configure(width=w, height=h, spp=spp, max_bounces=ref_cuda.max_bounces); ctx.camera_set((0, 0, -3))
    m = []
    for _ in range(6):
        ctx.frame_reset(); ctx.render_default(1); m.append(mape_score(gt, _img8(ctx.frame_download(), w, h)))
    mape_prod_default = float(np.mean(m))
    ctx.frame_reset(); ctx.render_default(32)                                # 1024 spp
    p1024 = ctx.frame_download()
    mape_converged = mape_score(gt, _img8(p1024, w, h))
    mean_rel = float(np.abs(p1024.mean(0) - np.nan_to_num(gt_f).mean(0)).max() / np.nan_to_num(gt_f).mean())

    # Expected SARSA, 8 training frames of 32 spp; the reference's frame 0 is NaN-poisoned (initial exclusive CDF), skip it there
    # The reference's own result varies from run to run (its TD update is a racy read-modify-write, radiance_volume.cu:283-301:
    # 0.524 .. 0.561 over five runs on one B200, the product 0.502 .. 0.506), so the reference side is the median of three runs.
    ref_runs = []
    for _ in range(3):
        ref_cuda.rmap_build()
        rmean, rlast, rst = ref_cuda.render_sarsa(9, 1)
        ref_runs.append(mape_score(gt, _img8(rmean, w, h)))
    mape_ref_sarsa = float(np.median(ref_runs))
    ref_sarsa_mpaths = w * h * spp / (rst[1:, 2].mean() * 1e-3) / 1e6
    ctx.frame_reset(); ctx.stats_reset(); ctx.radiance_map_build()
    ctx.render_sarsa(1); ctx.frame_reset(); ctx.stats_reset()                # same protocol: drop frame 0 from the image
    ctx.render_sarsa(8)
    mape_prod_sarsa = mape_score(gt, _img8(ctx.frame_download(), w, h))
    st_s = ctx.stats()
    out = dict(mape_ref_default_32spp=mape_ref_default, mape_prod_default_32spp=mape_prod_default, mape_refB_1024spp=mape_refb, mape_prod_default_1024spp=mape_converged,
               mean_rel_diff_1024spp=mean_rel, mape_ref_sarsa_8x32spp=mape_ref_sarsa, mape_ref_sarsa_runs=ref_runs, mape_prod_sarsa_8x32spp=mape_prod_sarsa,
               ref_kernels_default_mpaths_s=ref_default_mpaths, ref_kernels_sarsa_mpaths_s=ref_sarsa_mpaths,
               prod_sarsa_mpaths_s=st_s["paths"] / st_s["device_seconds"] / 1e6,
               ref_sarsa_avg_path_length_int_truncated=float(rst[1:, 0].mean()), prod_sarsa_avg_path_length=st_s["path_length_sum"] / st_s["paths"],
               ref_sarsa_nan_pixels_frame0=float(rst[0, 4]), ref_sarsa_nan_pixels_last=float(rst[-1, 4]))
    print("MAPE parity:", json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mape_parity.json"), "w"), indent=1)
    assert abs(mape_prod_default - mape_ref_default) <= TOL_DMAPE_32, out
    assert abs(mape_converged - mape_refb) <= TOL_DMAPE_1024, out
    assert mean_rel <= TOL_MEAN, out
    assert -TOL_DMAPE_SARSA_BETTER <= mape_prod_sarsa - mape_ref_sarsa <= TOL_DMAPE_SARSA_WORSE, out
    assert mape_prod_sarsa < mape_prod_default and mape_ref_sarsa < mape_ref_default, out      # 256 importance-sampled spp beat 32 uniform spp on both sides


# ---- the other three BASELINE.json scenes (configs[2], [3], [4]'s geometry): same protocol, lighter (these scenes cost the reference's
# brute-force kernels 5-30x more per path than Cornell). Ground truth = the reference's default tracer at GT x 32 spp.
#            name            GT frames, 32-spp frames, SARSA frames after one dropped, reference SARSA runs, tol default, tol SARSA
# Tolerances: default tracer |dMAPE| <= 0.01 + 5 % of the reference's value (values 0.04 .. 1.6; measured r2: 0.0002 .. 0.04, the largest on door_room,
# whose 32-spp frames score ~1.6 and scatter by ~0.03 each; three frames per side);
# Expected SARSA: product - reference in [-better, +worse]. The product's deviations lower its error (see above); on the scenes with far more
# volumes than visits per frame they matter more than in Cornell -- measured r2 (profiles/r2_mape_parity.json): door_room -0.004 (1.082 vs 1.086),
# archway -0.061 (0.264 vs 0.326), Medieval_House -0.020 (0.016 vs 0.036); "worse" = the reference's run-to-run spread.
SCENE_CASES = [("door_room_lit", 8, 3, 4, 2, 0.05, (0.04, 0.02)),
               ("archway", 8, 3, 4, 2, 0.05, (0.09, 0.012)),
               ("medieval_norm", 8, 3, 4, 2, 0.05, (0.03, 0.005))]


@pytest.mark.parametrize("name,gt_frames,n32,sarsa_frames,ref_runs_n,tol_default,tol_sarsa", SCENE_CASES)
def test_mape_parity_other_scenes(ctx, request, all_scenes, name, gt_frames, n32, sarsa_frames, ref_runs_n, tol_default, tol_sarsa):
    from checkers import mape_score
    cam, env = CONFIG_SCENES[name]
    ref = request.getfixturevalue("ref_cuda_env1" if env else "ref_cuda")
    w, h, spp = ref.width, ref.height, ref.spp
    s = all_scenes[name]
    ref.scene_arrays(s["sv"], s["srgb"], s["lv"], s["lrgb"]); ref.camera(*cam)
    gt_f, _ = ref.render_default(gt_frames); gt = _img8(gt_f, w, h)
    refb_f, _ = ref.render_default(gt_frames)
    mape_refb = mape_score(gt, _img8(refb_f, w, h))
    r32 = [ref.render_default(1) for _ in range(n32)]
    mape_ref_default = float(np.mean([mape_score(gt, _img8(x, w, h)) for x, _ in r32]))
    ref_default_mpaths = w * h * spp / (np.mean([t[0, 1] for _, t in r32]) * 1e-3) / 1e6
    ref_default_len = float(np.mean([t[0, 0] for _, t in r32]))

    load_scene(ctx, s)
    ctx.configure(width=w, height=h, spp=spp, max_bounces=ref.max_bounces, env_light=env); ctx.camera_set(cam)
    m = []
    for _ in range(n32):
        ctx.frame_reset(); ctx.render_default(1); m.append(mape_score(gt, _img8(ctx.frame_download(), w, h)))
    mape_prod_default = float(np.mean(m))
    ctx.frame_reset(); ctx.stats_reset(); ctx.render_default(gt_frames)
    pgt = ctx.frame_download(); st_d = ctx.stats()
    mape_prod_gt = mape_score(gt, _img8(pgt, w, h))
    mean_rel = float(np.abs(pgt.mean(0) - np.nan_to_num(gt_f).mean(0)).max() / max(float(np.nan_to_num(gt_f).mean()), 1e-9))

    ref_runs = []
    for _ in range(ref_runs_n):
        nv_ref = ref.rmap_build()
        rmean, rlast, rst = ref.render_sarsa(sarsa_frames + 1, 1)
        ref_runs.append(mape_score(gt, _img8(rmean, w, h)))
    mape_ref_sarsa = float(np.median(ref_runs))
    ref_sarsa_mpaths = w * h * spp / (rst[1:, 2].mean() * 1e-3) / 1e6
    ctx.frame_reset(); ctx.stats_reset(); nv = ctx.radiance_map_build()
    assert nv == nv_ref
    ctx.render_sarsa(1); ctx.frame_reset(); ctx.stats_reset()
    ctx.render_sarsa(sarsa_frames)
    mape_prod_sarsa = mape_score(gt, _img8(ctx.frame_download(), w, h)); st_s = ctx.stats()
    out = dict(scene=name, radiance_volumes=nv, gt_spp=gt_frames * spp, mape_ref_default_32spp=mape_ref_default, mape_prod_default_32spp=mape_prod_default,
               mape_refB_gt_spp=mape_refb, mape_prod_default_gt_spp=mape_prod_gt, mean_rel_diff_gt_spp=mean_rel,
               mape_ref_sarsa=mape_ref_sarsa, mape_ref_sarsa_runs=ref_runs, mape_prod_sarsa=mape_prod_sarsa, sarsa_spp=sarsa_frames * spp,
               ref_kernels_default_mpaths_s=ref_default_mpaths, ref_kernels_sarsa_mpaths_s=ref_sarsa_mpaths,
               prod_default_mpaths_s=st_d["paths"] / st_d["device_seconds"] / 1e6, prod_sarsa_mpaths_s=st_s["paths"] / st_s["device_seconds"] / 1e6,
               ref_default_avg_path_length_int_truncated=ref_default_len, prod_default_avg_path_length=st_d["path_length_sum"] / st_d["paths"],
               ref_sarsa_avg_path_length_int_truncated=float(rst[1:, 0].mean()), prod_sarsa_avg_path_length=st_s["path_length_sum"] / st_s["paths"],
               ref_sarsa_nan_pixels_last=float(rst[-1, 4]))
    print("MAPE parity:", json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mape_parity_%s.json" % name), "w"), indent=1)
    assert abs(mape_prod_default - mape_ref_default) <= 0.01 + tol_default * mape_ref_default, out
    assert abs(mape_prod_gt - mape_refb) <= 0.01 + tol_default * mape_refb, out
    assert mean_rel <= 2e-2, out
    assert -tol_sarsa[0] <= mape_prod_sarsa - mape_ref_sarsa <= tol_sarsa[1], out
