import os
import sys

import pytest

# no network / no pretrained file in the test environments: the perceptual loss runs on seeded random VGG features,
# identical on both sides of every parity comparison (losses/L1_plus_perceptualLoss.py raises without this opt-in)
os.environ.setdefault("MMH_VGG19_RANDOM", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _no_tf32():
    """The oracle is fp32: cuDNN / cuBLAS must not silently run it in TF32 on the GPU box (torch's default for convs)."""
    try:
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:
        pass


def pytest_configure(config):
    _no_tf32()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout; ignored when the plugin is absent)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
