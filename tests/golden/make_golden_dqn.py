#!/usr/bin/env python
"""tests/golden/make_golden_dqn.py -- DQN fixture from the reference's committed weights.
Radiance_Map_Data/cornell_12_12.model (a DyNet TextFileSaver dump of the reference's trained Neural-Q network for the
Cornell box) is parsed and stored as float32; query points are hit points of the golden Cornell ray batch; `q` is the
network function (N/dq_network.cu, N/fc_layer.cu) evaluated in float32 by oracle/checkers.py::dqn_forward_numpy --
DyNet itself is not in this image, so the network FUNCTION is what is pinned, not DyNet's kernels."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from checkers import Oracle, dqn_forward_numpy, dynet_text_load  # noqa: E402

params, k_in = dynet_text_load("/root/reference/Radiance_Map_Data/cornell_12_12.model")
z = np.load(os.path.join(HERE, "scenes.npz")); s = {k.split("/")[1]: z[k] for k in z.files if k.startswith("cornell/")}
vertices = np.concatenate([s["sv"].ravel(), s["lv"].ravel()])
assert k_in == len(vertices) == 342
h = np.load(os.path.join(HERE, "closest_hit.npz"))
orc = Oracle(); orc.scene_set(s["sv"], s["srgb"], s["lv"], s["lrgb"])
ty, ix, t, pos = orc.closest_hit(h["cornell/org"], h["cornell/dir"], 512, 0)
pos = pos[ty == 2][:1000].astype(np.float32)
q = dqn_forward_numpy(params, vertices, pos)
first_line = open("/root/reference/Radiance_Map_Data/cornell_12_12.model").readline().strip()
np.savez_compressed(os.path.join(HERE, "dqn_cornell.npz"), params=params, k_in=k_in, pos=pos, q=q, header0=first_line)
# the other two committed networks: the Cornell network trained without epsilon decay, and the door-room network
# (K = 342 as well: 36 surfaces + the door-room light quad, tests/golden/make_presets.py door_room_lit)
zp = np.load(os.path.join(HERE, "scene_presets.npz")); sd = {k.split("/")[1]: zp[k] for k in zp.files if k.startswith("door_room_lit/")}
for out_name, model, sc, hits in (("dqn_cornell_no_decay.npz", "cornell_no_decay.model", s, "cornell"), ("dqn_door_room.npz", "door_room_12_12.model", sd, "door_room")):
    pm, kk = dynet_text_load("/root/reference/Radiance_Map_Data/" + model)
    vv = np.concatenate([sc["sv"].ravel(), sc["lv"].ravel()])
    assert kk == len(vv) == 342
    o2 = Oracle(); o2.scene_set(sc["sv"], sc["srgb"], sc["lv"], sc["lrgb"])
    ty2, _, _, pos2 = o2.closest_hit(h[hits + "/org"], h[hits + "/dir"], 512, 0)
    pos2 = pos2[ty2 == 2][:1000].astype(np.float32)
    q2 = dqn_forward_numpy(pm, vv, pos2)
    np.savez_compressed(os.path.join(HERE, out_name), params=pm, k_in=kk, pos=pos2, q=q2)
    print(out_name, "q range", q2.min(), q2.max(), "zero rows", int((q2.max(1) == 0).sum()), os.path.getsize(os.path.join(HERE, out_name)) // 1024, "KiB")
print("params", len(params), "q range", q.min(), q.max(), "zero rows", int((q.max(1) == 0).sum()), os.path.getsize(os.path.join(HERE, "dqn_cornell.npz")) // 1024, "KiB")
