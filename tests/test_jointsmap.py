"""Part-map rasteriser (generate_jointsmap, data/generic_dataset.py:30-78): the integer restatement of OpenCV in the
oracle against the golden vectors produced with the real cv2 calls (and against cv2 itself when it is importable),
and the kernel body on the host emulation against the oracle, pixel-exact."""
import os

import numpy as np
import pytest
import torch

import hostemu
from mmhand_b200 import runtime
from oracle import jointsmap_ref as J

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jointsmap.npz")


def test_oracle_reproduces_cv2_golden_vectors():
    g = np.load(GOLD)
    for i in range(len(g["uv"])):
        got = J.generate_jointsmap(g["uv"][i], g["depth"][i], 256, 256)
        assert got.shape == (256, 256, 3) and got.dtype == np.float64
        assert np.array_equal(got[:, :, 0], g["maps"][i].astype(np.float64)), i
        assert np.array_equal(got[:, :, 0], got[:, :, 1]) and np.array_equal(got[:, :, 0], got[:, :, 2])


def test_oracle_known_answers():
    # one horizontal bone of length 40 at depth 1: an ellipse of half-axes (20, 5) around (120, 100), colour 160
    uv = np.full((21, 2), -1000.0)
    z = np.arange(21, dtype=np.float64) + 500.0
    uv[0], uv[17] = (100.0, 100.0), (140.0, 100.0)
    z[0] = z[17] = 1.0
    m = J.generate_jointsmap(uv, z, 256, 256)[:, :, 0]
    ys, xs = np.nonzero(m == 160)
    assert ys.min() == 95 and ys.max() == 105 and xs.min() == 100 and xs.max() == 140
    assert m[100, 120] == 160 and m[94, 120] == 0


def test_restated_opencv_matches_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for _ in range(300):
        c = (int(rng.randint(-20, 280)), int(rng.randint(-20, 280)))
        ax = (int(rng.randint(0, 150)), 5)
        ang = int(rng.randint(-180, 181))
        poly = cv2.ellipse2Poly(c, ax, ang, 0, 360, 1)
        mine = J.ellipse2poly(c, ax, ang)
        assert np.array_equal(poly, np.array(mine)), (c, ax, ang)
        img = np.zeros((256, 256, 3))
        cv2.fillConvexPoly(img, poly, [1.0] * 3)
        mask = J.fill_convex_poly_mask(256, 256, mine)
        assert np.array_equal(img[:, :, 0] > 0, mask), (c, ax, ang)
        for y in np.nonzero(mask.any(1))[0]:          # the kernel's span representation: one run per row
            xs = np.nonzero(mask[y])[0]
            assert xs[-1] - xs[0] + 1 == len(xs)


def test_kernel_body_matches_oracle_pixel_exactly():
    runtime._TEST_OPS = hostemu.ops()
    try:
        from mmhand_b200.rasterize import generate_jointsmap
        g = np.load(GOLD)
        got = generate_jointsmap(torch.from_numpy(g["uv"]), torch.from_numpy(g["depth"]), 256, 256,
                                 dtype=torch.uint8).numpy()
        assert got.shape == (len(g["uv"]), 256, 256)
        for i in range(len(got)):
            assert np.array_equal(got[i], g["maps"][i]), (i, int((got[i] != g["maps"][i]).sum()))
        full = generate_jointsmap(torch.from_numpy(g["uv"][:2]), torch.from_numpy(g["depth"][:2]), 256, 256).numpy()
        assert full.shape == (2, 256, 256, 3) and full.dtype == np.float64
        assert np.array_equal(full[0], J.generate_jointsmap(g["uv"][0], g["depth"][0], 256, 256))
        # ragged / empty / non-square
        assert generate_jointsmap(torch.zeros(0, 21, 2), torch.zeros(0, 21), 256, 256, dtype=torch.uint8).shape == (0, 256, 256)
        rng = np.random.RandomState(5)
        uv, z = rng.uniform(0, 120, size=(3, 21, 2)), rng.uniform(1, 9, size=(3, 21))
        small = generate_jointsmap(torch.from_numpy(uv), torch.from_numpy(z), 128, 96, dtype=torch.uint8).numpy()
        for i in range(3):
            assert np.array_equal(small[i].astype(np.float64), J.generate_jointsmap(uv[i], z[i], 128, 96)[:, :, 0])
    finally:
        runtime._TEST_OPS = None


def test_far_away_and_nan_joints_do_not_draw(monkeypatch):
    """A joint millions of pixels away (or NaN) drops its bones instead of walking millions of rows (kernel guard);
    everything else is drawn as the oracle draws it without those bones."""
    runtime._TEST_OPS = hostemu.ops()
    try:
        from mmhand_b200.rasterize import generate_jointsmap
        rng = np.random.RandomState(8)
        uv = rng.uniform(16, 240, size=(2, 21, 2))
        z = rng.uniform(200, 700, size=(2, 21))
        uv[0, 4] = (1e12, 3.0)                 # joint 4: only bone (3, 4)
        uv[1, 8] = (float("nan"), 50.0)        # joint 8: only bone (7, 8)
        got = generate_jointsmap(torch.from_numpy(uv), torch.from_numpy(z), 256, 256, dtype=torch.uint8).numpy()
        for i, dead in ((0, (3, 4)), (1, (7, 8))):
            monkeypatch.setattr(J, "BONES", tuple(b for b in J.BONES if b[0] != dead))
            clean = uv[i].copy()
            clean[dead[1]] = clean[dead[0]]
            want = J.generate_jointsmap(clean, z[i], 256, 256)[:, :, 0]
            monkeypatch.undo()
            assert np.array_equal(got[i].astype(np.float64), want), i
    finally:
        runtime._TEST_OPS = None
