"""Drop-in for the reference's ``data/mmhand_dataset_data_loader.py`` (:9-48): same class name, constructor (``opt``),
``len()`` and iteration, plus the ``set_epoch`` that ``train.py:53`` calls on it when ``--distributed`` and that the
reference class does not have (its training run dies with an AttributeError at the end of the first epoch).

For the RHD / STB datasets the batches are the COMPACT form of mmhand_b200.loader (uint8 frames + keypoints; heatmaps,
colour normalisation and depth decoding happen on the GPU inside ``MMHandModel.set_input`` with the reference's own
arithmetic, bit-identically -- tests/test_loader.py). ``MMH_COMPACT_INPUT=0`` or any other ``opt.dataset`` keeps the
reference's dataset classes (resolved from the reference tree) behind the same loader."""
import os

from mmhand_b200.loader import CompactHandDataset, DeviceInputLoader


class MMHandDatasetDataLoader(DeviceInputLoader):
    def __init__(self, opt):
        kind = getattr(opt, 'dataset', None)
        compact = os.environ.get("MMH_COMPACT_INPUT", "1") != "0" and kind in ('rhd', 'stb')
        if compact:
            dataset = CompactHandDataset(opt)
        elif kind == 'stb':
            from data.stb_dataset import STBdataset
            dataset = STBdataset(opt)
        elif kind == 'rhd':
            from data.rhd_dataset import RHDdataset
            dataset = RHDdataset(opt)
        else:
            from data.mmhand_dataset import MMHandDataset
            dataset = MMHandDataset(opt)
        super().__init__(opt, dataset=dataset)
        if not getattr(opt, 'distributed', False) or getattr(opt, 'local_rank', 0) == 0:
            print("dataset [%s] was created" % type(self.dataset).__name__)
