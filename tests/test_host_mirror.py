"""-m "not gpu": the C++ host mirror (host/rlpt_host.cpp: Scene::load_cornell_box_scene, load_custom_scene, Material,
AreaLight, normals) against the golden scenes the reference's own loaders produced (tests/golden/scenes.npz), bit for bit."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT, bits

LIB = os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200", "lib", "librlpt_host.so")
MODELS = "/root/reference/Models"


def load(path, lights_in_obj=False, preset=0):
    if not os.path.exists(LIB):
        pytest.skip("librlpt_host.so not built")
    L = ctypes.CDLL(LIB)
    ns, nl, nvert = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    p = path.encode() if path else None
    null = ctypes.c_void_p()
    rc = L.rlpt_host_load_scene(p, int(lights_in_obj), preset, ctypes.byref(ns), ctypes.byref(nl), *([null] * 9), ctypes.byref(nvert))
    assert rc == 0
    d = dict(sv=np.zeros((ns.value, 9), np.float32), srgb=np.zeros((ns.value, 3), np.float32), snrm=np.zeros((ns.value, 3), np.float32), slum=np.zeros(ns.value, np.float32),
             lv=np.zeros((nl.value, 9), np.float32), lrgb=np.zeros((nl.value, 3), np.float32), lnrm=np.zeros((nl.value, 3), np.float32), llum=np.zeros(nl.value, np.float32))
    vert = np.zeros(nvert.value, np.float32)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = L.rlpt_host_load_scene(p, int(lights_in_obj), preset, ctypes.byref(ns), ctypes.byref(nl), *[ptr(d[k]) for k in ("sv", "srgb", "snrm", "slum", "lv", "lrgb", "lnrm", "llum")], ptr(vert), ctypes.byref(nvert))
    assert rc == 0
    d["vertices"] = vert
    return d


def same(a, b):
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def test_cornell_box_matches_reference_loader(golden_scenes):
    d, g = load(None), golden_scenes["cornell"]
    assert len(d["sv"]) == 36 and len(d["lv"]) == 2 and len(d["vertices"]) == 342
    for k in ("sv", "srgb", "snrm", "slum", "lv", "lrgb", "lnrm", "llum"):
        assert same(d[k], g[k]), k
    assert same(d["vertices"], np.concatenate([g["sv"].ravel(), g["lv"].ravel()]))


@pytest.mark.parametrize("name,lights_in_obj", [("door_room", False), ("archway", False), ("simple_room", False), ("Medieval_House", False), ("complex_light_room", True)])
def test_obj_import_matches_reference_loader(golden_scenes, name, lights_in_obj):
    path = os.path.join(MODELS, name + ".obj")
    if not os.path.exists(path):
        pytest.skip("reference models not present")
    d, g = load(path, lights_in_obj), golden_scenes[name]
    keys = ("sv", "srgb", "snrm", "slum", "lv", "lrgb", "llum") + (() if lights_in_obj else ("lnrm",))
    for k in keys:
        assert same(d[k], g[k]), (name, k)


def test_presets_and_errors(tmp_path):
    path = os.path.join(MODELS, "door_room.obj")
    if not os.path.exists(path):
        pytest.skip("reference models not present")
    d = load(path, False, preset=1)                               # the door-room light quad commented in object_importer.cu:215-219
    assert len(d["lv"]) == 2 and np.allclose(d["lrgb"], 8.0) and np.allclose(d["srgb"][24:36], [0.75, 0.15, 0.15]) and np.allclose(d["srgb"][12:24], [0.15, 0.15, 0.75])
    n = load(os.path.join(MODELS, "Medieval_House.obj"), False, preset=2)
    assert len(n["lv"]) == 0 and np.abs(n["sv"]).max() <= 1.0 + 1e-5      # normalised into [-1, 1]
    L = ctypes.CDLL(LIB)
    null = ctypes.c_void_p()
    assert L.rlpt_host_load_scene(str(tmp_path / "missing.obj").encode(), 0, 0, *([null] * 13)) == 1
    # ragged input: quads, x/y/z indices, blank lines, unknown records, out-of-range index
    obj = tmp_path / "t.obj"
    obj.write_text("o thing\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\n\nvt 0 0\ns off\nf 1/1/1 2/2/1 3/3/1 4/4/1\nf 1 2 3\nf 1 2 9\n")
    t = load(str(obj), False, preset=2)
    assert len(t["sv"]) == 3


def saved_volume_surfaces(path, n_max=4096):
    if not os.path.exists(LIB):
        pytest.skip("librlpt_host.so not built")
    L = ctypes.CDLL(LIB)
    sv, rgb, nrm = np.zeros((n_max, 9), np.float32), np.zeros((n_max, 3), np.float32), np.zeros((n_max, 3), np.float32)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    n = L.rlpt_host_saved_volumes_to_surfaces(str(path).encode(), n_max, ptr(sv), ptr(rgb), ptr(nrm))
    return n, sv[:max(n, 0)], rgb[:max(n, 0)], nrm[:max(n, 0)]


def test_saved_radiance_volumes_as_geometry_match_reference(tmp_path):
    """RadianceVolume::read_radiance_volumes_to_surfaces (radiance_volume.cu:377-515: reader, get_vertices,
    build_surfaces) in the host mirror against the reference's own output on its committed selected_sarsa.txt, bit for bit."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "saved_volumes.npz"))
    f = tmp_path / "selected.txt"
    f.write_bytes(g["text"].tobytes())
    n, sv, rgb, nrm = saved_volume_surfaces(f)
    assert n == len(g["sv"]) == 3 * 288                      # 12 x 12 quads x 2 triangles per volume
    assert same(sv, g["sv"]) and same(rgb, g["rgb"]) and same(nrm, g["nrm"])
    # the hemisphere has diameter 0.15 around the volume, colours run green -> red with the distribution value
    first = np.array(f.read_text().splitlines()[0].split()[:3], np.float32)
    assert np.all(np.linalg.norm(sv[:288].reshape(-1, 3) - first, axis=1) <= 0.15 * 1.0001)
    assert np.allclose(rgb[:, 0] + rgb[:, 1], 1.0, atol=1e-6) and np.all(rgb[:, 2] == 0) and rgb[:, 0].max() == 1.0
    # a missing file reports failure (the reference prints and carries on); ragged lines are skipped
    assert saved_volume_surfaces(tmp_path / "nope.txt")[0] == -1
    (tmp_path / "ragged.txt").write_text("0 0 0 0 1 0 0.5 0.5\n" + f.read_text().splitlines()[1] + "\n")
    assert saved_volume_surfaces(tmp_path / "ragged.txt")[0] == 288
