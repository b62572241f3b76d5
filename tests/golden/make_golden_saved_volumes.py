#!/usr/bin/env python
"""Golden vectors for the saved-radiance-volume geometry (RENDER_SAVED_RADIANCE_VOLUMES): the reference's own
RadianceVolume::read_radiance_volumes_to_surfaces (oracle/_ref/libref_host.so = unmodified reference sources built for the host)
run on the reference's committed Radiance_Map_Data/selected_radiance_volumes/selected_sarsa.txt. Run in the build container:
    python tests/golden/make_golden_saved_volumes.py        -> tests/golden/saved_volumes.npz
The input text travels inside the fixture (the GPU box has no /root/reference)."""
import ctypes, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = "/root/reference/Radiance_Map_Data/selected_radiance_volumes/selected_sarsa.txt"
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_host.so"))
text = open(SRC).read()
n = 288 * len([l for l in text.splitlines() if l.strip()])
sv, rgb, nrm = np.zeros((n, 9), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
count = L.ref_saved_volumes_to_surfaces(SRC.encode(), n, p(sv), p(rgb), p(nrm))
assert count == n, (count, n)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "saved_volumes.npz"), text=np.frombuffer(text.encode(), np.uint8), sv=sv, rgb=rgb, nrm=nrm)
print("wrote saved_volumes.npz:", count, "surfaces")
