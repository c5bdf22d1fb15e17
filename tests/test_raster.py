"""Heatmap rasteriser: known answers of the oracle restatement and the kernel body on the host emulation."""
import os

import numpy as np
import torch

import hostemu
from mmhand_b200 import runtime
from oracle.raster_ref import get_heatmaps, get_heatmaps_batch


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "heatmaps_ref.npz")


def test_oracle_reproduces_reference_golden_vectors():
    """tests/golden/heatmaps_ref.npz = Genericdataset.get_heatmaps of the reference itself (oracle/make_golden_raster.py)."""
    g = np.load(GOLD)
    for i in range(len(g["uv"])):
        assert np.array_equal(get_heatmaps(g["uv"][i]), g["maps"][i]), i
    assert np.array_equal(get_heatmaps_batch(g["uv"]), g["maps"])


def test_kernel_body_reproduces_reference_golden_vectors():
    runtime._TEST_OPS = hostemu.ops()
    try:
        from mmhand_b200.rasterize import get_heatmaps as gpu_heatmaps
        g = np.load(GOLD)
        got = gpu_heatmaps(torch.from_numpy(g["uv"]), (256, 256)).numpy()
        assert np.array_equal(got, g["maps"])
    finally:
        runtime._TEST_OPS = None


def test_oracle_known_answers():
    m = get_heatmaps(np.array([[100.0, 60.0]]))[0]
    assert m.shape == (256, 256) and m.dtype == np.float32
    assert m[60, 100] == 1.0                                   # row = y, column = x
    assert abs(m[60, 103] - np.exp(-9.0 / 72.0)) < 1e-7
    assert (m > 0).sum() == 1041
    assert abs(m[m > 0].min() - 0.010508660465) < 1e-9
    ys, xs = np.nonzero(m)
    assert ((xs - 100) ** 2 + (ys - 60) ** 2).max() <= 332     # zero iff D2 > 332.2958...
    assert np.array_equal(get_heatmaps_batch(np.array([[[100.0, 60.0]]]))[0, 0], m)


def test_kernel_body_matches_oracle_bit_exactly():
    runtime._TEST_OPS = hostemu.ops()
    try:
        from mmhand_b200.rasterize import get_heatmaps as gpu_heatmaps
        rng = np.random.RandomState(49)
        uv = rng.uniform(-20, 276, size=(6, 21, 2))
        uv[0, 0] = (0.0, 0.0)
        uv[0, 1] = (255.0, 255.0)
        uv[0, 2] = (128.0, 128.0 + np.sqrt(332.2958775))       # threshold grazing
        uv[0, 3] = (1e12, -1e12)                               # far outside: must not overflow the bounding box
        uv[0, 4] = (-18.5, 100.0)                              # disc touches the frame from outside
        uv[0, 5] = (273.0, 273.9)
        uv[0, 6] = (-40.0, 300.0)
        uv[0, 7] = (float("nan"), 10.0)                        # NaN propagates like numpy's exp
        got = gpu_heatmaps(torch.from_numpy(uv), (256, 256)).numpy()
        want = get_heatmaps_batch(uv, (256, 256))
        assert np.array_equal(got, want, equal_nan=True)
        assert np.isnan(got[0, 7]).all() and not np.isnan(got[0, :7]).any()
        assert gpu_heatmaps(torch.zeros(0, 21, 2, dtype=torch.float64), (256, 256)).shape == (0, 21, 256, 256)
    finally:
        runtime._TEST_OPS = None


def test_cords_to_map_variant():
    """Offline pose maps of the dataset tool (tool/generate_pose_map_RHD.py:22-29): un-thresholded, HWC, MISSING_VALUE
    joints skipped -- bit-exact against the line-by-line restatement, on the host emulation of the kernel."""
    import hostemu
    from mmhand_b200 import runtime
    from mmhand_b200.rasterize import cords_to_map
    from oracle.raster_ref import cords_to_map as ref
    runtime._TEST_OPS = hostemu.ops()
    try:
        rng = np.random.RandomState(5)
        cords = rng.uniform(-10, 70, size=(3, 18, 2))
        cords[0, 3] = (-1, 20.5)                 # missing y
        cords[1, 7] = (12.25, -1)                # missing x
        cords[2, 0] = (30, 30)                   # integer joint: exactly 1.0 at its pixel
        cords[2, 1] = (500.0, -300.0)            # far outside: underflows to 0 / denormals
        got = cords_to_map(torch.from_numpy(cords), (48, 64)).numpy()
        assert got.shape == (3, 48, 64, 18) and got.dtype == np.float32
        for i in range(3):
            want = ref(cords[i], (48, 64))
            assert np.array_equal(got[i], want), (i, np.abs(got[i] - want).max())
        assert got[0, :, :, 3].max() == 0 and got[1, :, :, 7].max() == 0 and got[2, 30, 30, 0] == 1.0
        ints = np.array([[10, 20], [-1, 5], [40, 63]])       # integer annotations, as json.loads returns them
        assert np.array_equal(cords_to_map(torch.from_numpy(ints), (48, 64)).numpy(), ref(ints, (48, 64)))
    finally:
        runtime._TEST_OPS = None
