"""-m gpu: the C++ host mirror (reinforcement-light-rays-pathtracer_b200/host/: Scene, Camera, Renderer, RadianceMap, SDLScreen, NeuralQPathtracer,
PretrainedPathtracer, train_q_value_network, the NCCL hook) exercised on the device through lib/rlpt_example -- a program shaped like the
reference's main.cu -- and compared with the ctypes path the other tests use: the two drive the same C ABI, so their BMPs must be byte-identical."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_scene

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200", "lib", "rlpt_example")


def _run(args, cwd):
    if not os.path.exists(EXE):
        pytest.skip("lib/rlpt_example not built (needs NCCL headers)")
    r = subprocess.run([EXE] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


@pytest.mark.parametrize("method", [0, 1])
def test_example_program_matches_ctypes_path(ctx, golden_scenes, tmp_path, method):
    out = _run(["--scene", "cornell", "--method", method, "--frames", 3, "--spp", 8, "--size", 128, "--out", "cpp.bmp",
                "--save-q", "q.txt", "--save-vertices", "vertices.txt"], tmp_path)
    assert "wrote cpp.bmp" in out
    load_scene(ctx, golden_scenes["cornell"])
    ctx.configure(width=128, height=128, spp=8, max_bounces=80); ctx.camera_set((0, 0, -3))
    if method == 1:
        ctx.radiance_map_build()
        for _ in range(3):
            ctx.render_sarsa(1)
    else:
        ctx.render_default(3)
    ctx.frame_save_bmp(str(tmp_path / "py.bmp"))
    a, b = open(tmp_path / "cpp.bmp", "rb").read(), open(tmp_path / "py.bmp", "rb").read()
    assert len(a) == len(b) == 122 + 128 * 128 * 4
    if method == 0:
        assert a == b                                              # same Philox paths, same kernels: the same bytes
    else:                                                          # float atomics order the TD sums differently run to run: a few pixels may differ by one level
        da = np.frombuffer(a[122:], np.uint8).astype(int); db = np.frombuffer(b[122:], np.uint8).astype(int)
        assert np.mean(da == db) >= 0.98 and np.abs(da - db).max() <= 24
        v = np.loadtxt(tmp_path / "vertices.txt").ravel()
        s = golden_scenes["cornell"]
        assert len(v) == 342 and np.allclose(v, np.concatenate([s["sv"].ravel(), s["lv"].ravel()]), rtol=1e-5, atol=1e-6)      # ofstream << float: 6 significant digits
        first = open(tmp_path / "q.txt").readline().strip()
        assert first == "144" and sum(1 for _ in open(tmp_path / "q.txt")) == 1 + ctx.n_vol


def test_offline_trainer_and_neural_q_programs(ctx, golden_scenes, tmp_path):
    """NN_Q_Value_Trainer's flow: a Q table saved by the SARSA run is fitted by train_q_value_network; the saved DyNet text model is then rendered
    with by PretrainedPathtracer; NeuralQPathtracer trains online and saves its model after every frame."""
    _run(["--scene", "cornell", "--method", 1, "--frames", 4, "--spp", 8, "--size", 128, "--out", "sarsa.bmp", "--save-q", "q.txt", "--save-vertices", "vertices.txt"], tmp_path)
    out = _run(["--train-q", "q.txt", "vertices.txt", "fit.model", "--epochs", 3, "--batch", 128], tmp_path)
    line = [l for l in out.splitlines() if l.startswith("trained 3 epochs")][0]
    loss0, loss1 = float(line.split("loss ")[1].split(" ->")[0]), float(line.split("-> ")[1].split(",")[0])
    assert loss1 < loss0 and out.count("Loss:") == 3 and out.count("Error:") == 3
    load_scene(ctx, golden_scenes["cornell"])
    ctx.dqn_load_text(str(tmp_path / "fit.model"))                 # a well-formed DyNet text model for the Cornell vertex list
    q = ctx.dqn_forward(np.zeros((4, 3), np.float32))
    assert q.shape == (4, 144) and np.isfinite(q).all()
    out = _run(["--scene", "cornell", "--method", 4, "--model", "fit.model", "--frames", 1, "--spp", 4, "--size", 96, "--out", "pre.bmp"], tmp_path)
    assert os.path.getsize(tmp_path / "pre.bmp") == 122 + 96 * 96 * 4
    out = _run(["--scene", "cornell", "--method", 3, "--model", "nq.model", "--frames", 1, "--spp", 1, "--size", 64, "--batch", 1024, "--out", "nq.bmp"], tmp_path)
    assert "last loss" in out and os.path.exists(tmp_path / "nq.model") and os.path.getsize(tmp_path / "nq.bmp") == 122 + 64 * 64 * 4


def test_example_program_two_gpus_nccl_hook(tmp_path):
    """--gpus 2: one host thread per GPU, ncclCommInitAll, the Q accumulators all-reduced through host/nccl_hook.cpp every frame"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(["--scene", "cornell", "--method", 1, "--frames", 3, "--spp", 8, "--size", 128, "--gpus", 2, "--out", "two.bmp"], tmp_path)
    _run(["--scene", "cornell", "--method", 1, "--frames", 3, "--spp", 16, "--size", 128, "--gpus", 1, "--out", "one.bmp"], tmp_path)
    a = np.frombuffer(open(tmp_path / "two.bmp", "rb").read()[122:], np.uint8).astype(float)
    b = np.frombuffer(open(tmp_path / "one.bmp", "rb").read()[122:], np.uint8).astype(float)
    # 2 GPUs x 8 spp trace the samples 1 GPU x 16 spp traces (same Philox sample indices); the summed frame buffer carries the summed sample count
    ma, mb = a.reshape(-1, 4)[:, :3].mean(), b.reshape(-1, 4)[:, :3].mean()
    assert abs(ma - mb) <= 0.03 * mb, (ma, mb)
