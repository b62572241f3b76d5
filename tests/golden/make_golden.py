#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates the committed golden vectors from the REFERENCE ITSELF.

Runs only where /root/reference exists (this container): it drives oracle/_ref/libref_host.so -- the unmodified
reference engine compiled for the host by oracle/build_ref.sh -- and stores what it returns. Nothing here calls the
oracle restatement or the product. Outputs (all small):
  scenes.npz        triangles/colours/normals/luminances of the built-in Cornell box and of the bundled .obj models as
                    the reference's own loaders build them (scene.cu:8-60, object_importer.cu:8-412)
  closest_hit.npz   a fixed ray batch per scene with the reference's (type, index, t, position)   [host arithmetic]
  radiance_map.npz  Cornell radiance volumes (positions, surfaces), the flattened kd-tree, nearest-volume answers for
                    a fixed query batch, per-cell directions, and CDFs for four Q tables
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from checkers import Reference  # noqa: E402

MODELS = "/root/reference/Models"


def ray_batch(rs, R, scene, n_primary, cam):
    """jittered camera rays through a 512x512 image + one hemisphere-ish secondary ray from every surface hit"""
    px = rs.rand(n_primary, 2) * 512
    d = np.stack([px[:, 0] - 256, px[:, 1] - 256, np.full(n_primary, 512.0)], 1).astype(np.float32)
    o = np.tile(np.asarray(cam, np.float32), (n_primary, 1))
    ty, ix, t, pos = R.closest_hit(o, d)
    hit = ty == 2
    nrm = scene["snrm"][ix[hit]]
    d2 = rs.randn(int(hit.sum()), 3).astype(np.float32)
    d2 *= np.sign((d2 * nrm).sum(1, keepdims=True)).astype(np.float32)
    o2 = (pos[hit] + np.float32(1e-5) * d2).astype(np.float32)
    return np.concatenate([o, o2]), np.concatenate([d, d2])


def main():
    R = Reference("host")
    rs = np.random.RandomState(1984)
    scenes, hits = {}, {}
    cams = {"cornell": (0, 0, -3), "door_room": (0, 0.5, -0.9), "archway": (-1, 0.2, -0.99), "complex_light_room": (-1, -1, -0.4),
            "simple_room": (0, 0, -0.9), "Medieval_House": (0, 0, -3)}
    for name in ["cornell", "door_room", "archway", "complex_light_room", "simple_room", "Medieval_House"]:
        if name == "cornell":
            R.scene_cornell()
        else:
            R.scene_obj(os.path.join(MODELS, name + ".obj"), name == "complex_light_room")
        s = R.scene_get()
        for k, v in s.items():
            scenes[name + "/" + k] = v
        n = 8000 if name != "Medieval_House" else 4000
        o, d = ray_batch(rs, R, s, n, cams[name])
        ty, ix, t, pos = R.closest_hit(o, d)
        hits.update({name + "/org": o, name + "/dir": d, name + "/type": ty, name + "/index": ix, name + "/t": t})
        print(name, len(s["sv"]), "surfaces", len(s["lv"]), "lights;", len(o), "rays; types", np.bincount(ty, minlength=3))
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **scenes)
    np.savez_compressed(os.path.join(HERE, "closest_hit.npz"), **hits)

    R.scene_cornell()
    nv = R.rmap_build()
    pos, nrm, surf = R.rmap_volumes()
    tree = R.rmap_tree()
    q0, cdf0, vis0, irr0 = R.rmap_state()
    nq = 20000
    idx = rs.randint(0, nv, nq)
    qpos = (pos[idx] + rs.randn(nq, 3).astype(np.float32) * np.float32(0.02)).astype(np.float32)
    qnrm = nrm[idx].copy()
    qnrm[::50] = nrm[rs.randint(0, nv, len(qnrm[::50]))]          # some queries with a different surface's normal
    found = R.find_closest(qpos, qnrm)
    sub = rs.choice(nv, 256, replace=False)
    qs = {"constant": np.full((nv, 144), np.float32(100.0 / 144.0), np.float32),
          "one_hot": np.full((nv, 144), np.float32(0.8 / 144.0), np.float32),
          "lognormal": np.exp(rs.randn(nv, 144) * 2).astype(np.float32)}
    qs["one_hot"][np.arange(nv), rs.randint(0, 144, nv)] = 5.0
    cdfs = {}
    for k, q in qs.items():
        R.rmap_set_q(q)
        R.rmap_update_distributions()
        cdfs[k] = R.rmap_state()[1][sub].copy()
    gx, gy = np.meshgrid(np.arange(12) + 0.5, np.arange(12) + 0.5, indexing="ij")
    centre_dirs = np.stack([R.grid_dir(int(v), gx.ravel(), gy.ravel()) for v in sub[:16]])
    rgx, rgy = (rs.rand(512) * 12).astype(np.float32), (rs.rand(512) * 12).astype(np.float32)
    rand_dirs = np.stack([R.grid_dir(int(v), rgx, rgy) for v in sub[:16]])
    np.savez_compressed(os.path.join(HERE, "radiance_map.npz"), n_volumes=nv, pos=pos, surface=surf,
                        tree_dim=tree["dim"].astype(np.int8), tree_leaf=tree["leaf"].astype(np.int8), tree_left=tree["left"], tree_right=tree["right"],
                        tree_data=tree["data"], initial_cdf_row=cdf0[1], initial_irradiance=irr0[1:9],
                        query_pos=qpos, query_nrm=qnrm, query_found=found, cdf_volumes=sub,
                        q_lognormal_rows=qs["lognormal"][sub], q_one_hot_rows=qs["one_hot"][sub],
                        cdf_constant=cdfs["constant"], cdf_one_hot=cdfs["one_hot"], cdf_lognormal=cdfs["lognormal"],
                        centre_dirs=centre_dirs, rand_gx=rgx, rand_gy=rgy, rand_dirs=rand_dirs)
    for f in ("scenes.npz", "closest_hit.npz", "radiance_map.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
