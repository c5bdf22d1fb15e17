"""aug.py's image write-out (aug.py:57-71) on the host emulation of mmh_image_pack_bgr8 against the reference's own
chain: numpy de-normalisation, cv2.cvtColor(RGB2BGR), cv2.imwrite (PNG, lossless) and what comes back from disk."""
import numpy as np
import pytest
import torch

import hostemu
from mmhand_b200 import runtime


def test_bgr8_matches_the_reference_write_out(tmp_path):
    cv2 = pytest.importorskip("cv2")
    runtime._TEST_OPS = hostemu.ops()
    try:
        from mmhand_b200.augment import images_to_bgr8
        g = torch.Generator().manual_seed(7)
        fake = torch.tanh(torch.randn(3, 3, 40, 56, generator=g) * 2)
        fake[0, :, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, 1.5, -1.5, 0.00392157, -0.00392157, 0.5])    # ties, saturation
        k = torch.arange(0, 24).float()
        fake[1, 0, 1, :24] = (k + 0.5) / 127.5 - 1.0                # values that land on x.5 before rounding
        got = images_to_bgr8(fake).numpy()
        assert got.shape == (3, 40, 56, 3) and got.dtype == np.uint8
        for i in range(3):
            ref = fake[i].permute(1, 2, 0).numpy()                   # aug.py:57-60
            ref = (ref * 0.5 + 0.5) * 255.
            ref = cv2.cvtColor(ref, cv2.COLOR_RGB2BGR)
            path = str(tmp_path / ("img%d.png" % i))
            cv2.imwrite(path, ref)
            back = cv2.imread(path, cv2.IMREAD_COLOR)
            assert np.array_equal(got[i], back), (i, int((got[i] != back).sum()))
        assert images_to_bgr8(torch.zeros(0, 3, 8, 8)).shape == (0, 8, 8, 3)
    finally:
        runtime._TEST_OPS = None


def test_image_writer_and_compact_batches(tmp_path):
    """ImageWriter: batches submitted asynchronously end up on disk with exactly the bytes of aug.py's chain;
    generate_batch accepts the loader's compact batches."""
    cv2 = pytest.importorskip("cv2")
    runtime._TEST_OPS = hostemu.ops()
    try:
        from mmhand_b200.augment import ImageWriter, images_to_bgr8
        g = torch.Generator().manual_seed(3)
        batches = [torch.tanh(torch.randn(4, 3, 24, 32, generator=g) * 2) for _ in range(5)]
        with ImageWriter(workers=3, depth=2) as w:
            for bi, fake in enumerate(batches):
                w.submit(images_to_bgr8(fake), [str(tmp_path / ("b%d" % bi) / ("%d.png" % i)) for i in range(4)])
        for bi, fake in enumerate(batches):
            for i in range(4):
                ref = fake[i].permute(1, 2, 0).numpy()
                ref = cv2.cvtColor((ref * 0.5 + 0.5) * 255., cv2.COLOR_RGB2BGR)
                rp = str(tmp_path / "ref.png")
                cv2.imwrite(rp, ref)
                assert np.array_equal(cv2.imread(str(tmp_path / ("b%d" % bi) / ("%d.png" % i))), cv2.imread(rp))
        with pytest.raises(Exception):
            with ImageWriter(workers=1, depth=1) as w:
                w.submit(images_to_bgr8(batches[0]), ["/proc/nonexistent/x.png"] * 4)
    finally:
        runtime._TEST_OPS = None
