"""-m "not gpu": the oracle restatement against the reference ITSELF (oracle/_ref/libref_host*.so = the unmodified
reference sources compiled for the host by oracle/build_ref.sh). Skipped where the prebuilt library is absent.
Deterministic functions are compared bit for bit on fresh random inputs (beyond the committed golden vectors); the path
tracers use different random streams (the reference: XORWOW per pixel; the oracle: counter-based Philox) and are
compared statistically with the tolerance written at the assert.
"""
import os

import numpy as np
import pytest

from conftest import bits, load_scene

MODELS = "/root/reference/Models"


def _ref_scene(ref, name):
    if name == "cornell":
        ref.scene_cornell()
    else:
        if not os.path.exists(os.path.join(MODELS, name + ".obj")):
            pytest.skip("reference models not present")
        ref.scene_obj(os.path.join(MODELS, name + ".obj"), name == "complex_light_room")
    return ref.scene_get()


@pytest.mark.parametrize("name", ["cornell", "door_room", "archway", "Medieval_House"])
def test_golden_scenes_are_what_the_reference_loads(ref_host, golden_scenes, name):
    s = _ref_scene(ref_host, name)
    for k in ("sv", "srgb", "snrm", "slum", "lv", "lrgb", "lnrm", "llum"):
        assert np.array_equal(bits(s[k]), bits(golden_scenes[name][k])), k


@pytest.mark.parametrize("name", ["cornell", "archway", "Medieval_House"])
def test_closest_hit_bit_exact_on_fresh_rays(ref_host, oracle, name):
    s = _ref_scene(ref_host, name)
    load_scene(oracle, s)
    rs = np.random.RandomState(hash(name) % 1000)
    n = 20000 if name != "Medieval_House" else 3000
    lo, hi = s["sv"].reshape(-1, 3).min(0), s["sv"].reshape(-1, 3).max(0)
    org = (lo + (hi - lo) * rs.rand(n, 3)).astype(np.float32)
    dir = rs.randn(n, 3).astype(np.float32)
    dir[::7] *= np.float32(1e-3)                                   # unnormalised inputs: Ray::Ray normalises
    rty, rix, rt, rpos = ref_host.closest_hit(org, dir)
    oty, oix, ot, opos = oracle.closest_hit(org, dir, ref_host.height, 0)
    assert np.array_equal(rty, oty) and np.array_equal(rix, oix) and np.array_equal(bits(rt), bits(ot))
    hit = rty != 0
    assert np.array_equal(bits(rpos[hit]), bits(opos[hit]))


def test_radiance_map_and_nearest_volume_bit_exact(ref_host, oracle):
    s = _ref_scene(ref_host, "door_room")
    load_scene(oracle, s)
    nv = ref_host.rmap_build()
    assert oracle.rmap_build() == nv and nv > 30000
    rp, rn, rsf = ref_host.rmap_volumes(); op, on, osf = oracle.rmap_volumes()
    assert np.array_equal(bits(rp), bits(op)) and np.array_equal(bits(rn), bits(on)) and np.array_equal(rsf, osf)
    rt, ot = ref_host.rmap_tree(), oracle.rmap_tree()
    for k in ("dim", "leaf", "left", "right"):
        assert np.array_equal(rt[k], ot[k]), k
    assert np.array_equal(bits(rt["data"]), bits(ot["data"]))
    rs = np.random.RandomState(4)
    idx = rs.randint(0, nv, 30000)
    pos = (rp[idx] + rs.randn(len(idx), 3).astype(np.float32) * np.float32(0.02)).astype(np.float32)
    pos[::101] = rs.uniform(-2, 2, (len(pos[::101]), 3))
    nrm = rn[idx].copy(); nrm[::53] = rn[rs.randint(0, nv, len(nrm[::53]))]
    assert np.array_equal(ref_host.find_closest(pos, nrm), oracle.find_closest(pos, nrm, 0))


def test_cdf_matches_reference(ref_host, oracle):
    s = _ref_scene(ref_host, "cornell")
    load_scene(oracle, s)
    nv = ref_host.rmap_build(); assert oracle.rmap_build() == nv
    q = np.exp(np.random.RandomState(1984).randn(nv, 144) * 2).astype(np.float32)
    q[::5] = np.float32(0.8 / 144)
    ref_host.rmap_set_q(q); ref_host.rmap_update_distributions()
    oracle.rmap_set_q(q); oracle.rmap_update_distributions()
    rcdf, ocdf = ref_host.rmap_state()[1], oracle.rmap_state()[1]
    assert (np.abs(rcdf - ocdf) / np.maximum(np.abs(rcdf), 1e-30)).max() <= 1e-5       # north_star bar


def test_default_render_statistics(oracle):
    """method 0 at 512x512x2spp: mean radiance within 3%, mean path length within 2% (different RNG streams; the standard
    error of the image mean over 524k paths is about 0.5%)"""
    from checkers import Reference
    if not Reference.available("host", "_spp2"):
        pytest.skip("oracle/_ref/libref_host_spp2.so not built")
    ref = Reference("host", "_spp2")
    s = _ref_scene(ref, "cornell")
    ref.camera(0.0, 0.0, -3.0)
    load_scene(oracle, s)
    rimg, rstats = ref.render_default(1)
    oimg, ost = oracle.render_frame(0, ref.width, ref.height, ref.spp, cam=(0, 0, -3), fma_mode=0)
    oimg = oimg / ref.spp
    rimg = np.nan_to_num(rimg)
    assert abs(float(oimg.mean()) - float(rimg.mean())) <= 0.03 * float(rimg.mean())
    # per-channel means: colour bleeding from the red/blue walls must agree too
    assert np.allclose(oimg.mean(0), rimg.mean(0), rtol=0.04)
    # coarse image agreement: 16x16 block means, relative error of the block grid
    def blocks(a):
        return a.reshape(ref.width // 32, 32, ref.height // 32, 32, 3).mean((1, 3))
    rb, ob = blocks(rimg.reshape(ref.width, ref.height, 3)), blocks(oimg.reshape(ref.width, ref.height, 3))
    assert np.abs(rb - ob).mean() <= 0.08 * rb.mean()
