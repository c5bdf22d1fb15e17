"""ORACLE support (test infrastructure): import the *real* reference modules from /root/reference with the
minimal shims of SURVEY.md section 8c (stub apex / skimage, random-init VGG19, CPU no-ops for .cuda()).
Only usable where /root/reference exists (the build container); never on the GPU box and never by the product.
"""
import importlib
import os
import sys
import types

REF = os.environ.get("MMH_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _RefImport:
    """Context manager: reference tree first on sys.path, our same-named packages hidden."""

    NAMES = ("models", "losses", "util", "options", "data")

    def __enter__(self):
        self.saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in self.NAMES}
        for k in self.saved:
            del sys.modules[k]
        sys.path.insert(0, REF)
        return self

    def __exit__(self, *a):
        sys.path.remove(REF)
        self.ref_modules = {k: v for k, v in sys.modules.items() if k.split(".")[0] in self.NAMES}
        for k in self.ref_modules:
            del sys.modules[k]
        sys.modules.update(self.saved)


def load_reference_nets():
    """Returns (Generator, Discriminator, network_utils module, ImagePool, L1_plus_perceptualLoss)."""
    with _RefImport():
        gen = importlib.import_module("models.Generator")
        dis = importlib.import_module("models.Discriminator")
        nu = importlib.import_module("models.network_utils")
        pool = importlib.import_module("util.image_pool")
        l1p = importlib.import_module("losses.L1_plus_perceptualLoss")
    return gen.Generator, dis.Discriminator, nu, pool.ImagePool, l1p.L1_plus_perceptualLoss


def load_reference_dataset_module():
    """data.generic_dataset of the reference (Genericdataset.get_heatmaps / gen_heatmap / gaussian_kernel,
    generate_jointsmap), importable here with three shims: ``from cv2 import cv2`` (removed from OpenCV >= 4.6: alias
    cv2.cv2 = cv2), ``easydict`` (stub, only used by the dataset constructors) and ``np.math`` (removed in numpy 2: alias
    of the math module). The arithmetic itself runs unmodified."""
    import math

    import cv2
    import numpy as np
    cv2.cv2 = cv2
    sys.modules.setdefault("cv2.cv2", cv2)
    if "easydict" not in sys.modules:
        _stub("easydict", EasyDict=dict)
    if not hasattr(np, "math"):
        np.math = math
    with _RefImport():
        gd = importlib.import_module("data.generic_dataset")
    return gd


def load_reference_dataset_classes():
    """(RHDdataset, STBdataset, MMHandDatasetDataLoader) of the reference (data/rhd_dataset.py, data/stb_dataset.py,
    data/mmhand_dataset_data_loader.py), importable here with the shims of load_reference_dataset_module plus
    ``np.bool`` (removed in numpy 1.24; generic_dataset.py:111 uses it)."""
    import numpy as np
    load_reference_dataset_module()
    if not hasattr(np, "bool"):
        np.bool = bool
    with _RefImport():
        rhd = importlib.import_module("data.rhd_dataset")
        stb = importlib.import_module("data.stb_dataset")
        ldr = importlib.import_module("data.mmhand_dataset_data_loader")
    return rhd.RHDdataset, stb.STBdataset, ldr.MMHandDatasetDataLoader


def load_reference_model_class():
    """models.MMHandModel.MMHandModel of the reference, importable on CPU."""
    import torch
    import torchvision

    amp = _stub("apex.amp")
    par = _stub("apex.parallel", DistributedDataParallel=object, convert_syncbn_model=lambda m: m)
    _stub("apex", amp=amp, parallel=par)
    _stub("skimage.draw", circle=None, line_aa=None, polygon=None)
    _stub("skimage", draw=sys.modules["skimage.draw"])
    _orig_vgg = torchvision.models.vgg19

    def vgg19(pretrained=False, **kw):
        return _orig_vgg(weights=None)

    torchvision.models.vgg19 = vgg19
    with _RefImport():
        mm = importlib.import_module("models.MMHandModel")
    if not torch.cuda.is_available():
        class _Cuda:
            @staticmethod
            def is_available():
                return True
        mm.torch = types.SimpleNamespace(**{k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})
        mm.torch.cuda = _Cuda
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.nn.DataParallel = lambda m, device_ids=None: m
    return mm.MMHandModel


from mmhand_b200.options import make_opt  # noqa: E402,F401  (kept under its historical name for the tests)
