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
print("params", len(params), "q range", q.min(), q.max(), "zero rows", int((q.max(1) == 0).sum()), os.path.getsize(os.path.join(HERE, "dqn_cornell.npz")) // 1024, "KiB")
