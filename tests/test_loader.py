"""Device-side input pipeline: CompactHandDataset / DeviceInputLoader against the reference's own RHDdataset
on a tiny synthetic dataset written to disk -- same pairing, and after ``set_input`` bit-identical model inputs (colour
normalisation, depth decoding, heatmaps), on the host emulation of the kernels."""
import os
import random
import types

import numpy as np
import pytest
import torch

import hostemu
from mmhand_b200 import runtime
from mmhand_b200.loader import CompactHandDataset, DeviceInputLoader, compact_from_reference
from mmhand_b200.options import make_opt
from oracle import ref_shims
from synth_dataset import make_rhd


@pytest.fixture
def data_dir():
    """pytest's tmp_path contains 'test', which the reference's datasets take for the evaluation split
    (generic_dataset.py:114)."""
    import shutil
    import tempfile
    d = tempfile.mkdtemp(prefix="mmh_rhd_")
    yield d
    shutil.rmtree(d, ignore_errors=True)


@pytest.fixture
def emu():
    runtime._TEST_OPS = hostemu.ops(f32=True)
    yield
    runtime._TEST_OPS = None


def _opt(root, **kw):
    d = dict(dataroot=root, dataset='rhd', augmentation_ratio=0.5, isTrain=True, batchSize=2, nThreads=0,
             distributed=False, max_dataset_size=float("inf"), seed=49, local_rank=0)
    d.update(kw)
    return types.SimpleNamespace(**d)


def _model(S):
    from models.MMHandModel import MMHandModel
    torch.manual_seed(1)
    m = MMHandModel(make_opt(batchSize=2, fineSize=S, ngf=16, ndf=16, pool_size=0, local_rank='cpu', seed=3))
    m.master = False
    return m


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
def test_compact_batches_equal_the_reference_loader(data_dir, emu):
    S = 32
    root = make_rhd(os.path.join(data_dir, "rhd"), n=8, size=S)
    RHD, _, RefLoader = ref_shims.load_reference_dataset_classes()
    random.seed(11)
    ref_loader = RefLoader(_opt(root))
    random.seed(11)
    mine = DeviceInputLoader(_opt(root))
    assert len(mine) == len(ref_loader) == 4
    assert mine.dataset.image_source == ref_loader.dataset.image_source          # same split, same shuffled pairing
    assert mine.dataset.image_target == ref_loader.dataset.image_target
    mine.set_epoch(3)                                                            # train.py:53; a no-op when not distributed
    m_ref, m_mine, m_c = _model(S), _model(S), _model(S)
    n = 0
    for rb, cb in zip(ref_loader, mine):
        assert rb['H1_path'] == cb['H1_path'] and rb['H2_path'] == cb['H2_path']
        assert set(cb) >= {'H1_u8', 'H2_u8', 'D1_u8', 'D2_u8', 'P1_uv', 'P2_uv', 'C1', 'C2'}
        assert cb['H1_u8'].dtype == torch.uint8 and cb['H1_u8'].is_pinned() == torch.cuda.is_available()
        m_ref.set_input(rb)                                   # the reference's fp32 / fp64 tensors
        m_mine.set_input(cb)                                  # compact form: everything computed by the kernels
        m_c.set_input(compact_from_reference(rb))             # the reference's batch minus the heatmaps (C1 / C2 keypoints)
        for k in ('input_H1', 'input_H2', 'input_D1', 'input_D2', 'input_P1', 'input_P2'):
            a = getattr(m_ref, k)
            assert torch.equal(a, getattr(m_mine, k)), k
            assert torch.equal(a, getattr(m_c, k)), k
        assert torch.equal(rb['C1'], cb['C1']) and torch.equal(rb['C2'], cb['C2'])
        # callers that index the reference's keys (aug.py:43-47) get them computed on demand, bit-identically
        assert 'P1' not in cb and 'H1' not in cb
        for k in ('H1', 'H2', 'D1', 'D2', 'P1', 'P2'):
            assert torch.equal(cb[k].cpu(), rb[k].float()), k
        # and the oracle's restatement of the dataset arithmetic (used by the GPU tests, where the reference tree is
        # absent) is the reference's
        from oracle.raster_ref import decode_depth_u8, normalize_image_u8
        assert torch.equal(torch.from_numpy(normalize_image_u8(cb['H1_u8'].numpy(), bgr=True)), rb['H1'])
        assert torch.equal(torch.from_numpy(decode_depth_u8(cb['D2_u8'].numpy())), rb['D2'].float())
        n += 1
    assert n == 2
    bytes_ref = sum(v.numel() * v.element_size() for v in rb.values() if isinstance(v, torch.Tensor))
    bytes_mine = sum(v.numel() * v.element_size() for v in cb.values() if isinstance(v, torch.Tensor))
    assert bytes_mine * 10 < bytes_ref


def test_eval_split_and_max_dataset_size(data_dir):
    root = make_rhd(os.path.join(data_dir, "rhd"), n=10, size=16)
    random.seed(2)
    ds = CompactHandDataset(_opt(root, isTrain=False, augmentation_ratio=0.3))
    assert len(ds) == 7 and sorted(ds.image_source) == ds.image_target      # first 1 - ratio share, shuffled sources
    assert [os.path.basename(p) for p in ds.image_target] == ["%05d.png" % i for i in range(7)]
    ld = DeviceInputLoader(_opt(root, batchSize=1, max_dataset_size=3))
    assert len(ld) == 3 and sum(1 for _ in ld) == 3
    it = ds[0]
    assert it['H1_u8'].shape == (16, 16, 3) and it['P1_uv'].shape == (21, 2) and it['P1_uv'].dtype == torch.float64
