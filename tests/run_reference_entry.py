"""Test harness: execute an UNMODIFIED entry point of the reference (``train.py`` / ``aug.py``) with this repository in
front of the reference tree on sys.path -- ``models.*``, ``losses.*``, ``util.image_pool`` resolve to this repository,
``options.*``, ``data.*``, ``util.visualizer`` / ``util.util`` to the reference.

  python tests/run_reference_entry.py train <train.py arguments...>
  python tests/run_reference_entry.py aug <ckp> <dataroot> <DST> <dataset> <ratio> <device>

What is shimmed is the ENVIRONMENT, not the scripts: packages the reference imports that are not installed here
(easydict, skimage.draw, dominate, visdom), API names that newer numpy / OpenCV dropped (np.bool, np.math, cv2.cv2),
two option attributes aug.py's hand-built ``opt`` lacks and -- only when there is no CUDA device -- the
CUDA-only calls of options/base_options.py (set_device, the nccl backend) plus the host emulation of the kernels.
Needs the reference tree: build container only."""
import math
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MMH_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_environment_shims():
    import cv2
    import numpy as np
    import torch
    cv2.cv2 = cv2
    sys.modules.setdefault("cv2.cv2", cv2)
    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(np, "math"):
        np.math = math

    class EasyDict(dict):                      # the attribute-dict behaviour aug.py relies on
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                # aug.py's opt lacks the two attributes the loader reads
                if k == "max_dataset_size":
                    return float("inf")
                if k == "seed":
                    return 49
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    if "easydict" not in sys.modules:
        _stub("easydict", EasyDict=EasyDict)
    _stub("skimage.draw", circle=None, line_aa=None, polygon=None)
    _stub("skimage", draw=sys.modules["skimage.draw"])
    tags = _stub("dominate.tags", **{k: None for k in ("meta", "h3", "table", "tr", "td", "p", "a", "img", "br")})
    _stub("dominate", tags=tags)
    _stub("visdom")
    os.environ.setdefault("MMH_VGG19_RANDOM", "1")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29800 + os.getpid() % 100))
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    if not torch.cuda.is_available():
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import hostemu
        from mmhand_b200 import runtime
        runtime._TEST_OPS = hostemu.ops(f32=True)
        torch.cuda.set_device = lambda *a, **k: None
        _init = torch.distributed.init_process_group

        def init_process_group(backend=None, **kw):
            return _init(backend="gloo", **kw)

        torch.distributed.init_process_group = init_process_group
        _to = torch.Tensor.to

        def to(self, *a, **k):                 # aug.py: .to(<int device>) on a box without CUDA
            if a and isinstance(a[0], int):
                a = ("cpu",) + a[1:]
            return _to(self, *a, **k)

        torch.Tensor.to = to
        _mto = torch.nn.Module.to
        torch.nn.Module.to = lambda self, *a, **k: _mto(self, *(("cpu",) + a[1:] if a and isinstance(a[0], int) else a), **k)


if __name__ == "__main__":
    which = sys.argv[1]
    sys.path[:0] = [ROOT, REF]
    install_environment_shims()
    script = os.path.join(REF, {"train": "train.py", "aug": "aug.py"}[which])
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")
