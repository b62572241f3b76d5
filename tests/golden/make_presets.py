#!/usr/bin/env python
"""tests/golden/make_presets.py -- scene presets for BASELINE.json configs[2] and [3] that the reference only has as
commented-out code, so its own loader build (make_golden.py) cannot produce them:
  door_room_lit    door_room.obj with the door-room light quad I, J, K, L (G/objects/object_importer.cu:215-219,233-237,
                   power 8) and the door-room colours (:153-155,161-163)
  medieval_norm    Medieval_House.obj with the commented normalisation `scale = 2 / max_difference` (:119) and no lights
                   (lit by ENVIRONMENT_LIGHT)
Built with the product's C++ host mirror (host/rlpt_host.cpp, ImportPreset), which test_host_mirror.py checks bit for bit
against the reference's loader on the presets the reference does build. Runs only where /root/reference exists."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_host_mirror import MODELS, load  # noqa: E402

out = {}
for name, obj, preset in [("door_room_lit", "door_room.obj", 1), ("medieval_norm", "Medieval_House.obj", 2)]:
    d = load(os.path.join(MODELS, obj), False, preset=preset)
    for k in ("sv", "srgb", "lv", "lrgb"):
        out[name + "/" + k] = d[k]
    print(name, len(d["sv"]), "surfaces", len(d["lv"]), "lights", "bounds", d["sv"].reshape(-1, 3).min(0), d["sv"].reshape(-1, 3).max(0))
np.savez_compressed(os.path.join(HERE, "scene_presets.npz"), **out)
