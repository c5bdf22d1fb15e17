"""Device-side input pipeline: loaders that hand ``MMHandModel.set_input`` the *compact* form of a batch.

The reference's dataset workers build every sample on the CPU (data/generic_dataset.py:133-180): two colour frames
normalised to fp32, two depth frames decoded and normalised to float64 x 3 channels, and 2 x 21 Gaussian heatmaps of
256 x 256 fp32 (35 ms per pose and core) -- 14 MB per sample cross the bus, 78 % of them heatmaps. At 300+ samples/s
per GPU that is > 20 CPU cores and 4 GB/s per GPU. Here the workers only read and decode the files; keypoints, uint8
colour frames and uint8 depth frames travel (0.8 MB per sample) and ``set_input`` rasterises / normalises them on the
device with the reference's own arithmetic (mmh_heatmap_rasterize, mmh_image_unpack_u8, mmh_depth_unpack_u8).

  CompactHandDataset      the reference's Genericdataset / RHDdataset / STBdataset over the same on-disk layout
                          (annotation.pickle, <root>/<folder>/<image>, depth frame = path with "color" -> "depth"), same
                          pairing (``_get_src_tgt``: sort, split by augmentation_ratio, shuffled sources), compact items
  DeviceInputLoader       MMHandDatasetDataLoader's API (``__len__``, ``__iter__``, max_dataset_size) plus the ``set_epoch``
                          that train.py:53 calls and the reference class lacks
  compact_from_reference  turns a batch of the reference's own loader into the compact form where possible (its
                          'C1' / 'C2' entries carry the keypoints: the 2 x 21 heatmaps need not be copied)
"""
import os
import pickle
import random

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset


def compact_from_reference(batch):
    """Batch dict of the reference's loader -> the same batch with 'P1' / 'P2' replaced by 'P1_uv' / 'P2_uv' (taken
    from 'C1' / 'C2' = [B, 21, (u, v, z)], data/generic_dataset.py:161-167). Everything else is passed through."""
    if 'C1' not in batch or 'C2' not in batch:
        return batch
    out = {k: v for k, v in batch.items() if k not in ('P1', 'P2')}
    out['P1_uv'] = batch['C1'][..., :2].to(torch.float64).contiguous()
    out['P2_uv'] = batch['C2'][..., :2].to(torch.float64).contiguous()
    return out


class CompactBatch(dict):
    """A compact batch that still answers the reference's keys: ``batch['P1']``, ``['H1']``, ``['D1']`` ... are computed
    on first access -- on the GPU, by the kernels ``MMHandModel.set_input`` uses -- so that callers written against the
    reference's loader (``aug.py:43-47`` indexes the six tensors directly) run unchanged. ``'P1' in batch`` stays False:
    ``set_input`` sees the compact form and never materialises the fp32 tensors on the host side."""

    def __missing__(self, key):
        from . import runtime
        from .rasterize import get_heatmaps
        if key in ('P1', 'P2') and (key + '_uv') in self:
            ref = self['H1_u8'] if 'H1_u8' in self else self['H1']
            hw = tuple(ref.shape[1:3]) if ref.dtype == torch.uint8 else tuple(ref.shape[2:])
            return get_heatmaps(self[key + '_uv'], hw, sigma=float(self.get('sigma', 6.0)))
        if key in ('H1', 'H2', 'D1', 'D2') and (key + '_u8') in self:
            ops = runtime.get_ops(None)
            src = self[key + '_u8'].to(ops.device, non_blocking=True)
            out = torch.empty(src.shape[0], 3, src.shape[1], src.shape[2], dtype=torch.float32, device=ops.device)
            if key[0] == 'H':
                ops.image_unpack_u8(src, out, swap_rb=bool(self.get('u8_bgr', False)))
            else:
                ops.depth_unpack_u8(src, out, hi=1, lo=2, div=float(self.get('depth_div', 700.0)))
            return out
        raise KeyError(key)


class CompactHandDataset(Dataset):
    """opt: dataroot, dataset ('rhd' | 'stb'), augmentation_ratio, isTrain -- the options the reference's datasets read."""

    def __init__(self, opt):
        super().__init__()
        import cv2
        self.cv2 = cv2
        self.opt = opt
        self.root_dir = opt.dataroot
        with open(os.path.join(self.root_dir, "annotation.pickle"), "rb") as handle:
            self.annotations = pickle.load(handle)
        images = []
        kind = getattr(opt, 'dataset', 'rhd')
        for folder in self.annotations.keys():
            for image in self.annotations[folder].keys():
                path = os.path.join(self.root_dir, folder, image)
                if kind == 'stb':                       # data/stb_dataset.py:24-33: SK colour frames only
                    camera, spec, _ = image.split('_')
                    if camera != 'BB' and spec == 'color':
                        images.append(path)
                elif folder == 'color':                 # data/rhd_dataset.py:27-32
                    images.append(path)
        if kind == 'stb':
            def sort_fn(x):                             # data/stb_dataset.py:35-40
                *_, folder, name = x.split('/')
                return int(folder[1]), folder[2], int(name[0:-4].split('_')[-1])
        else:
            def sort_fn(x):                             # data/rhd_dataset.py:34-37
                return int(x.split('/')[-1][0:-4])
        self.image_source, self.image_target = self._get_src_tgt(opt.augmentation_ratio, images, sort_fn)

    def _get_src_tgt(self, ratio, data, sort_fn):
        """data/generic_dataset.py:96-128: the last `ratio` share trains, the first 1 - ratio share is augmented."""
        assert len(data) > 0
        data.sort(key=sort_fn)
        sep = int((1 - ratio) * len(data))
        if 'test' in self.root_dir:
            assert not self.opt.isTrain
            tgt = data
        else:
            tgt = data[sep:] if self.opt.isTrain else data[:sep]
        src = tgt.copy()
        random.shuffle(src)
        return src, tgt

    def __len__(self):
        return len(self.image_source)

    def get_labels(self, image_path):                   # data/generic_dataset.py:201-206
        *_, folder, name = image_path.split('/')
        if "joints" in name:
            name = name.split('_')
            name = name[0] + "_" + name[1] + "_" + name[-1]
        return self.annotations[folder][name]

    def __getitem__(self, item):
        cv2 = self.cv2
        h_1, h_2 = self.image_source[item], self.image_target[item]
        a1, a2 = self.get_labels(h_1), self.get_labels(h_2)
        batch = {}
        for k, path, ann in (('1', h_1, a1), ('2', h_2, a2)):
            batch['H%s_u8' % k] = torch.from_numpy(cv2.imread(path))                             # BGR, HWC uint8
            batch['D%s_u8' % k] = torch.from_numpy(cv2.imread(path.replace("color", "depth")))
            uv = np.array(ann['uv_coord'], dtype=np.float64)
            z = np.expand_dims(np.array(ann['depth']), -1) / 700.0 * 255
            batch['P%s_uv' % k] = torch.from_numpy(uv)
            batch['C%s' % k] = torch.tensor(np.concatenate([uv, z], axis=-1))
        batch['u8_bgr'] = True
        batch['H1_path'], batch['H2_path'] = h_1, h_2
        return batch


class DeviceInputLoader():
    """Drop-in for data/mmhand_dataset_data_loader.py::MMHandDatasetDataLoader over CompactHandDataset (or over any
    dataset passed as ``dataset``): same constructor argument (``opt``), ``len()``, iteration and max_dataset_size cut,
    plus ``set_epoch`` (train.py:53 calls it when distributed; the reference class has none)."""

    def __init__(self, opt, dataset=None):
        self.opt = opt
        self.dataset = dataset if dataset is not None else CompactHandDataset(opt)
        self.distributed_sampler = None
        seed = getattr(opt, 'seed', 0)
        init_fn = None
        if getattr(opt, 'distributed', False):
            self.distributed_sampler = torch.utils.data.distributed.DistributedSampler(self.dataset)
            init_fn = lambda w_id: np.random.seed(seed)
        collate = self._collate if isinstance(self.dataset, CompactHandDataset) else None
        self.dataloader = DataLoader(self.dataset, batch_size=opt.batchSize, shuffle=False, pin_memory=True,
                                     sampler=self.distributed_sampler, worker_init_fn=init_fn,
                                     num_workers=int(getattr(opt, 'nThreads', 0)), collate_fn=collate)

    @staticmethod
    def _collate(items):
        out = CompactBatch()
        for k in items[0]:
            v = [it[k] for it in items]
            if isinstance(v[0], torch.Tensor):
                out[k] = torch.stack(v)
            elif isinstance(v[0], bool):
                out[k] = v[0]
            else:
                out[k] = v
        return out

    def set_epoch(self, epoch):
        if self.distributed_sampler is not None:
            self.distributed_sampler.set_epoch(epoch)

    def __len__(self):
        return min(len(self.dataset), getattr(self.opt, 'max_dataset_size', float("inf")))

    def __iter__(self):
        limit = getattr(self.opt, 'max_dataset_size', float("inf"))
        for i, data in enumerate(self.dataloader):
            if i >= limit:
                break
            yield data
