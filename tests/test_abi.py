"""The C-ABI shared library loads and exports every entry point include/mmhand_sm100.h declares (no compute)."""
import os
import re

from mmhand_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmhand_sm100.h")).read()
    return sorted(set(re.findall(r"\b(mmh_[a-z0-9_]+)\s*\(", src)))


def test_cuda_library_exports_header():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = L.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert lib.mmh_is_device_build() == 1
    assert lib.act_bytes == 2
    assert lib.mmh_version() >= 100


def test_python_binding_covers_header():
    assert set(_declared()) - {"mmh_act_bytes"} <= set(L.EXPORTS)


def test_hostemu_exports_same_abi():
    import hostemu
    lib = hostemu.load()
    for n in _declared():
        assert hasattr(lib, n), n
    assert lib.mmh_is_device_build() == 0


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    with pytest.raises(L.MmhError):
        L.load(str(tmp_path / "nope.so"))


def test_no_cpu_fallback_in_runtime():
    import pytest
    import torch
    from mmhand_b200 import runtime
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    assert runtime._TEST_OPS is None
    with pytest.raises(RuntimeError):
        runtime.get_ops()
