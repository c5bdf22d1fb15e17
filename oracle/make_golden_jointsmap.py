"""Seeded + adversarial poses of the part-map golden vectors (``poses``). tests/golden/jointsmap.npz itself is written
by oracle/make_golden_raster.py from the reference's own ``generate_jointsmap``; running this file writes the same
maps from ``generate_jointsmap_cv2`` (the reference's lines on the real OpenCV) -- they are identical.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import jointsmap_ref as J  # noqa: E402


def poses(n=24, seed=49):
    rng = np.random.RandomState(seed)
    uv = rng.uniform(16, 240, size=(n, 21, 2))
    z = rng.uniform(200, 700, size=(n, 21))
    uv[1] = np.round(uv[1])                                   # integer pixels: exact 0 / 45 / 90 degree bones occur
    uv[2, :, 1] = 100.0                                       # all bones horizontal
    uv[3, :, 0] = 77.0                                        # all bones vertical
    uv[4] = np.round(uv[4]); uv[4, 1] = uv[4, 0] + (13, 13); uv[4, 5] = uv[4, 0] + (-20, 20)      # diagonals
    uv[5, 6] = uv[5, 5]                                       # zero-length bone
    z[6] = np.round(z[6] / 100) * 100                         # equal depths between bones
    uv[7] = rng.uniform(-60, 320, size=(21, 2))               # joints outside the frame (clipLine)
    uv[8] = rng.uniform(-300, 600, size=(21, 2))
    uv[9, :, 0] = rng.uniform(-5, 5, size=21)                 # hugging the left border
    uv[10, :, 1] = rng.uniform(250, 262, size=21)             # hugging the bottom border
    return uv, z


if __name__ == "__main__":
    import cv2
    uv, z = poses()
    maps = np.stack([J.generate_jointsmap_cv2(uv[i], z[i], 256, 256)[:, :, 0].astype(np.uint8) for i in range(len(uv))])
    out = os.path.join(ROOT, "tests", "golden", "jointsmap.npz")
    np.savez_compressed(out, uv=uv, depth=z, maps=maps, cv2_version=cv2.__version__)
    print("wrote", out, maps.shape, "non-zero pixels per pose:", (maps > 0).reshape(len(uv), -1).sum(1).tolist())
