"""Process-wide access to the kernel library: one ``Ops`` per CUDA device.

There is deliberately no CPU implementation behind this: without the CUDA library or without a CUDA device
``get_ops`` raises. (``_TEST_OPS`` lets the CPU unit tests of the host logic inject the host-emulation library;
nothing in the package sets it.)
"""
import torch

from .kernels import Ops

_OPS = {}
_TEST_OPS = None


def get_ops(device=None) -> Ops:
    if _TEST_OPS is not None:
        return _TEST_OPS
    if device is None or torch.device(device).type != "cuda":
        if not torch.cuda.is_available():
            raise RuntimeError("mmhand_b200 runs on CUDA (sm_100a) only: no CUDA device is visible and there is no "
                               "CPU fallback")
        idx = torch.cuda.current_device()
    else:
        d = torch.device(device)
        idx = d.index if d.index is not None else torch.cuda.current_device()
    if idx not in _OPS:
        _OPS[idx] = Ops.cuda(idx)
    return _OPS[idx]


class World:
    """Data-parallel group (one process per GPU, torch.distributed)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.size > 1 else 0

    def all_reduce(self, t):
        if self.size > 1:
            self.dist.all_reduce(t)
