"""Test helper: a tiny RHD-shaped dataset on disk in the layout the reference's datasets read
(data/rhd_dataset.py:15-44, data/generic_dataset.py:80-95,133-180): ``annotation.pickle`` =
{'color': {name: {'uv_coord': [[u, v] x 21], 'depth': [z x 21]}}, 'depth': {name: {}}}, frames <root>/color/<name> and
<root>/depth/<name> (depth value = 256 * G + R of the PNG)."""
import os
import pickle

import numpy as np


def make_rhd(root, n=6, size=64, seed=3):
    import cv2
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "color"), exist_ok=True)
    os.makedirs(os.path.join(root, "depth"), exist_ok=True)
    ann = {"color": {}, "depth": {}}
    for i in range(n):
        name = "%05d.png" % i
        img = rng.randint(0, 256, size=(size, size, 3)).astype(np.uint8)
        img = cv2.blur(img, (5, 5))
        d = rng.randint(200, 700, size=(size, size))
        d = cv2.blur(d.astype(np.float32), (7, 7)).astype(np.int32)
        dep = np.zeros((size, size, 3), np.uint8)
        dep[:, :, 1] = d // 256          # G
        dep[:, :, 2] = d % 256           # R
        cv2.imwrite(os.path.join(root, "color", name), img)
        cv2.imwrite(os.path.join(root, "depth", name), dep)
        ann["color"][name] = {"uv_coord": rng.uniform(4, size - 4, size=(21, 2)).tolist(),
                              "depth": rng.uniform(200, 700, size=21).tolist()}
        ann["depth"][name] = {}
    with open(os.path.join(root, "annotation.pickle"), "wb") as f:
        pickle.dump(ann, f)
    return root
