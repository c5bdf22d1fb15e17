"""Golden vectors of the two rasterisers produced by the REFERENCE'S OWN functions (data/generic_dataset.py imported
from /root/reference with the shims of oracle/ref_shims.py::load_reference_dataset_module):

  tests/golden/heatmaps_ref.npz   Genericdataset.get_heatmaps(uv, (256, 256), 6) for seeded + adversarial poses
  tests/golden/jointsmap.npz      generate_jointsmap(uv, depth, 256, 256) (channel 0 as uint8; the three are equal)

Run in the build container (the reference does not travel to the GPU box):  python oracle/make_golden_raster.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from oracle.make_golden_jointsmap import poses as jm_poses  # noqa: E402


def heatmap_poses(n=8, seed=49):
    rng = np.random.RandomState(seed)
    uv = rng.uniform(16, 240, size=(n, 21, 2))
    uv[0] = np.array([[10.0 * j, 7.0 * j] for j in range(21)])                       # integer pixels
    uv[1, :4] = np.array([[0.0, 0.0], [255.0, 255.0], [-30.0, 40.0], [300.0, 128.0]])    # borders, outside
    r = np.sqrt(332.2958775)
    uv[2] = np.array([[128.0 + r * np.cos(t), 128.0 + r * np.sin(t)] for t in np.linspace(0, 6.2, 21)])  # threshold grazing
    uv[3, :3] = np.array([[-18.5, 100.0], [273.0, 273.9], [-40.0, 300.0]])
    return uv


if __name__ == "__main__":
    import cv2
    gd = ref_shims.load_reference_dataset_module()
    ds = gd.Genericdataset.__new__(gd.Genericdataset)          # the three methods used do not touch instance state
    uv = heatmap_poses()
    maps = np.stack([gd.Genericdataset.get_heatmaps(ds, uv[i], (256, 256), 6).numpy() for i in range(len(uv))])
    assert maps.dtype == np.float32 and maps.shape == (len(uv), 21, 256, 256)
    out = os.path.join(ROOT, "tests", "golden", "heatmaps_ref.npz")
    np.savez_compressed(out, uv=uv, maps=maps, numpy_version=np.__version__)
    print("wrote", out, maps.shape, "non-zero:", int((maps > 0).sum()))

    juv, jz = jm_poses()
    jm = np.stack([gd.generate_jointsmap(juv[i], jz[i], 256, 256) for i in range(len(juv))])
    assert jm.dtype == np.float64 and jm.shape == (len(juv), 256, 256, 3)
    assert np.array_equal(jm[..., 0], jm[..., 1]) and np.array_equal(jm[..., 0], jm[..., 2])
    out = os.path.join(ROOT, "tests", "golden", "jointsmap.npz")
    np.savez_compressed(out, uv=juv, depth=jz, maps=jm[..., 0].astype(np.uint8), cv2_version=cv2.__version__,
                        source="reference data/generic_dataset.py::generate_jointsmap")
    print("wrote", out, jm.shape)
