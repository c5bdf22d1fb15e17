"""ctypes binding of libmmhand_sm100.so (the C ABI declared in include/mmhand_sm100.h).

The product path has exactly one backend: the CUDA library built in-tree by ``__graft_entry__.build()``.
If it is missing, importing a kernel raises -- there is no CPU fallback. (CPU tests of the host logic load
the host-emulation build explicitly through ``tests/hostemu.py``; nothing in this package does.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmhand_sm100.so")
MAX_TAPS = 64


class MmhError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_int64), ("a_ld", C.c_int32), ("C", C.c_int32),
        ("w", C.c_void_p), ("T", C.c_int32), ("N", C.c_int32),
        ("shift", C.c_int32 * MAX_TAPS),
        ("w_slot", C.c_int32 * MAX_TAPS), ("w_taps", C.c_int32),
        ("M", C.c_int64),
        ("Hg", C.c_int32), ("Wg", C.c_int32), ("Hv", C.c_int32), ("Wv", C.c_int32),
        ("out", C.c_void_p), ("out_f32", C.c_int32), ("out_ld", C.c_int32),
        ("out_img_rows", C.c_int64),
        ("out_wg", C.c_int32), ("out_sh", C.c_int32), ("out_sw", C.c_int32),
        ("out_h0", C.c_int32), ("out_w0", C.c_int32),
        ("zero_invalid", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32), ("n_store", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_int64), ("a_ld", C.c_int32), ("C", C.c_int32),
        ("dy", C.c_void_p), ("M", C.c_int64), ("dy_ld", C.c_int32), ("N", C.c_int32),
        ("T", C.c_int32), ("shift", C.c_int32 * MAX_TAPS),
        ("dw", C.c_void_p), ("tap_index", C.c_int32 * MAX_TAPS), ("dw_taps", C.c_int32),
        ("N_store", C.c_int32), ("C_store", C.c_int32), ("split_k", C.c_int32),
        ("dbg_lbo_sbo_swap", C.c_int32),
    ]


_lib = None


def load(path=None):
    """Load the shared library (once). ``path`` overrides the in-tree CUDA build (tests only)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise MmhError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the mmhand_b200 kernels)" % p)
    lib = C.CDLL(p)
    lib.mmh_last_error.restype = C.c_char_p
    lib.mmh_version.restype = C.c_int
    lib.mmh_is_device_build.restype = C.c_int
    _declare(lib)
    if path is None:
        _lib = lib
    return lib


def _declare(lib):
    vp = C.c_void_p
    lib.mmh_conv_plan_create.argtypes = [C.POINTER(ConvDesc), C.POINTER(vp)]
    lib.mmh_conv_plan_destroy.argtypes = [vp]
    lib.mmh_conv_run.argtypes = [vp, vp]
    lib.mmh_wgrad_plan_create.argtypes = [C.POINTER(WgradDesc), C.POINTER(vp)]
    lib.mmh_wgrad_plan_destroy.argtypes = [vp]
    lib.mmh_wgrad_run.argtypes = [vp, vp]
    for name, args in _SIMPLE_SIGS.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.argtypes = args
            fn.restype = C.c_int


_SIMPLE_SIGS = {}


def check(lib, rc):
    if rc != 0:
        raise MmhError(lib.mmh_last_error().decode("utf-8", "replace"))


def set_error_check(lib):
    return lambda rc: check(lib, rc)
