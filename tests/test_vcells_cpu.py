"""Host logic of the nearest-volume candidate cells (rlpt_radiance_host.cpp host_build_vcells + rlpt_device.cuh vcell_find):
whenever the cells decide a query, the answer is the reference kd search's (RadianceMap::find_closest_radiance_volume_iterative,
G/radiance_volumes/radiance_map.cu:150-203, restated by kd_find). Runs the product's own functions instantiated for the host
(tests/devfn_host.cpp); the GPU parity tests repeat the comparison on the device against the oracle and the reference."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="module")
def devfn():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libdevfn_host.so")
    srcs = [os.path.join(ROOT, "tests", "devfn_host.cpp"), os.path.join(CSRC, "rlpt_radiance_host.cpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs + [os.path.join(CSRC, "rlpt_device.cuh")]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I", CSRC, "-o", so] + srcs)
    lib = ctypes.CDLL(so)
    lib.devfn_check_vcells.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("scene,jitter,min_decided", [("cornell", 0.0, 0.999), ("cornell", 1e-6, 0.999), ("cornell", 0.02, 0.05), ("door_room", 0.0, 0.999), ("archway", 1e-6, 0.999)])
def test_candidate_cells_agree_with_kd_search(devfn, golden_scenes, scene, jitter, min_decided):
    sv = np.ascontiguousarray(golden_scenes[scene]["sv"], dtype=np.float32).reshape(-1, 9)
    out = np.zeros(9, dtype=np.float64)
    nq = 200000
    rc = devfn.devfn_check_vcells(sv.ctypes.data, len(sv), 0.001, 0.003, 0.4, nq, 7, jitter, out.ctypes.data)
    assert rc == 0
    mism, decided, nv, keys, listed, slots, decided2, xkeys, xlisted = out
    print(scene, jitter, dict(mismatches=mism, decided=decided / nq, by_second_level=decided2 / nq, volumes=nv, keys=keys, mean_list=listed / max(keys, 1), slots=slots,
                              xkeys=xkeys, mean_xlist=xlisted / max(xkeys, 1)))
    assert mism == 0
    assert decided / nq >= min_decided


@pytest.mark.parametrize("scene,cam", [("cornell", (0.0, 0.0, -3.0)), ("door_room", (0.0, 0.5, -0.9)), ("archway", (-1.0, 0.2, -0.99))])
def test_scan_unit_pretest_is_conservative(devfn, golden_scenes, scene, cam):
    """unit_candidates (the brute-force scan's pre-test over triangles / parallelogram pairs) never discards a primitive that
    the exact solve with the reference's arithmetic (tri_solve <- G/rays/ray.cu:39-74,115-141) accepts."""
    s = golden_scenes[scene]
    verts = np.ascontiguousarray(np.concatenate([np.asarray(s["sv"], np.float32).reshape(-1, 9), np.asarray(s["lv"], np.float32).reshape(-1, 9)]))
    camv = np.asarray(cam, np.float32)
    out = np.zeros(5, dtype=np.float64)
    devfn.devfn_check_units.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    nr = 400000
    assert devfn.devfn_check_units(verts.ctypes.data, len(verts), nr, 11, 512.0, camv.ctypes.data, out.ctypes.data) == 0
    missed, marked, accepted, units, pairs = out
    print(scene, dict(missed=missed, marked_per_ray=marked, accepted_per_ray=accepted, units=units, pairs=pairs))
    assert missed == 0
    assert accepted > 0.5 and marked < 6.0 * accepted + 2.0        # the pre-test is tight: besides the hits it keeps lines through a triangle behind the origin (single triangles) and the plane the ray starts on
