"""Launch defaults of data-parallel groups (mmhand_b200/runtime.py::World.tame_launches): programmatic dependent launch
goes off through the C ABI (mmh_set_pdl) unless MMH_PDL is set explicitly -- the combination with the peer-memory SyncBN
exchange stopped 8-GPU runs. Host logic only; the CUDA side is exercised by tests/test_gpu_ddp.py."""
import types

import torch

import hostemu
from mmhand_b200 import runtime


class _FakeLib:
    def __init__(self):
        self.calls = []

    def mmh_set_pdl(self, v):
        self.calls.append(v)
        return 0


def _world(size):
    w = runtime.World.__new__(runtime.World)
    w.size, w.rank, w.peer, w.seq, w._lib = size, 0, None, 0, None
    return w


def test_groups_switch_dependent_launches_off(monkeypatch):
    monkeypatch.delenv("MMH_PDL", raising=False)
    ops = types.SimpleNamespace(lib=_FakeLib(), device=torch.device("cuda", 0))
    _world(8).tame_launches(ops)
    assert ops.lib.calls == [0]
    # single process, or a CPU (gloo / host-emulation) group: untouched
    ops.lib.calls.clear()
    _world(1).tame_launches(ops)
    _world(4).tame_launches(types.SimpleNamespace(lib=ops.lib, device=torch.device("cpu")))
    assert ops.lib.calls == []
    # an explicit MMH_PDL wins
    monkeypatch.setenv("MMH_PDL", "1")
    _world(8).tame_launches(ops)
    assert ops.lib.calls == []


def test_pdl_switch_of_the_library_roundtrip():
    lib = hostemu.load()
    base = lib.mmh_get_pdl()
    assert lib.mmh_set_pdl(0) == 0 and lib.mmh_get_pdl() == 0
    assert lib.mmh_set_pdl(1) == 0 and lib.mmh_get_pdl() == 1
    assert lib.mmh_set_pdl(-1) == 0 and lib.mmh_get_pdl() == base
