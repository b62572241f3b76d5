"""Shared fixtures. `-m "not gpu"` covers the oracle against the golden vectors / the reference host build, the host
logic and the ABI surface; `-m gpu` holds the parity tests proper, which call the CUDA library through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200"))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


@pytest.fixture(scope="session")
def golden_scenes():
    z = np.load(os.path.join(GOLDEN, "scenes.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {k.split("/")[1]: z[k] for k in z.files if k.startswith(n + "/")} for n in names}


@pytest.fixture(scope="session")
def all_scenes(golden_scenes):
    """the bundled scenes plus the two presets of BASELINE.json configs[2] / [3] (tests/golden/make_presets.py): door_room_lit, medieval_norm"""
    z = np.load(os.path.join(GOLDEN, "scene_presets.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    out = dict(golden_scenes)
    out.update({n: {k.split("/")[1]: z[k] for k in z.files if k.startswith(n + "/")} for n in names})
    return out


# camera and ENVIRONMENT_LIGHT of the BASELINE.json configurations (G/main.cu:100-104; medieval: SURVEY section 7 preset)
CONFIG_SCENES = {"cornell": ((0.0, 0.0, -3.0), 0.0), "door_room_lit": ((0.0, 0.5, -0.9), 0.0), "archway": ((-1.0, 0.2, -0.99), 0.0), "medieval_norm": ((0.0, 0.0, -3.0), 1.0)}


@pytest.fixture(scope="session")
def golden_hits():
    z = np.load(os.path.join(GOLDEN, "closest_hit.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {k.split("/")[1]: z[k] for k in z.files if k.startswith(n + "/")} for n in names}


@pytest.fixture(scope="session")
def golden_rmap():
    return dict(np.load(os.path.join(GOLDEN, "radiance_map.npz")))


@pytest.fixture(scope="session")
def oracle():
    from checkers import Oracle, build_oracle
    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def ref_host():
    from checkers import Reference
    if not Reference.available("host"):
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference)")
    return Reference("host")


@pytest.fixture(scope="session")
def ref_cuda():
    from checkers import Reference
    if not Reference.available("cuda"):
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    return Reference("cuda")


@pytest.fixture(scope="session")
def ref_cuda_env1():
    """the reference's kernels built with ENVIRONMENT_LIGHT 1 (a compile-time constant there): oracle/build_ref.sh cuda 512 512 32 _env1 with RLPT_ORACLE_ENV=1.0f"""
    from checkers import Reference
    if not Reference.available("cuda", "_env1"):
        pytest.skip("oracle/_ref/libref_cuda_env1.so not built")
    return Reference("cuda", "_env1")


@pytest.fixture()
def ctx():
    import rlpt
    c = rlpt.Context(0)
    yield c
    c.close()


def load_scene(ctx_or_oracle, s):
    if hasattr(ctx_or_oracle, "scene_upload"):
        ctx_or_oracle.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"])
    else:
        ctx_or_oracle.scene_set(s["sv"], s["srgb"], s["lv"], s["lrgb"])
