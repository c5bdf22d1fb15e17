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
        ("bn_sums", C.c_void_p), ("bn_C", C.c_int32), ("reserved0", C.c_int32),
        ("bs_x", C.c_void_p), ("bs_coef", C.c_void_p), ("bs_save", C.c_void_p), ("bs_sums", C.c_void_p),
        ("bs_x_ld", C.c_int32), ("bs_xHg", C.c_int32), ("bs_xWg", C.c_int32), ("bs_H", C.c_int32), ("bs_W", C.c_int32),
        ("bs_pad", C.c_int32), ("bs_C", C.c_int32), ("bs_relu", C.c_int32), ("bs_dropout", C.c_int32),
        ("bs_drop_key", C.c_uint32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_int64), ("a_ld", C.c_int32), ("C", C.c_int32),
        ("dy", C.c_void_p), ("M", C.c_int64), ("dy_ld", C.c_int32), ("N", C.c_int32),
        ("T", C.c_int32), ("shift", C.c_int32 * MAX_TAPS),
        ("dw", C.c_void_p), ("tap_index", C.c_int32 * MAX_TAPS), ("dw_taps", C.c_int32),
        ("N_store", C.c_int32), ("C_store", C.c_int32), ("split_k", C.c_int32),
    ]


class ParamJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("s_n", C.c_int64), ("s_c", C.c_int64),
                ("kind", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("T", C.c_int32), ("Np", C.c_int32),
                ("Cp", C.c_int32), ("tile_begin", C.c_int32), ("reserved", C.c_int32)]


class Lay(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "H", "W", "Hg", "Wg", "h0", "w0", "phase", "ld", "c0", "C", "reserved")]


class NormAct(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("sl", Lay), ("coef", C.c_void_p),
        ("relu", C.c_int32), ("dropout", C.c_int32), ("drop_key", C.c_uint32), ("reserved", C.c_int32),
        ("resid", C.c_void_p), ("dst", C.c_void_p), ("dl", Lay),
        ("pad_lo", C.c_int32), ("pad_hi", C.c_int32), ("reflect", C.c_int32), ("reserved2", C.c_int32),
        ("dst_f32", C.c_void_p),
    ]


class GateFwd(C.Structure):
    _fields_ = [
        ("c1", C.c_void_p), ("x2o", C.c_void_p), ("x3o", C.c_void_p), ("sl", Lay), ("coef", C.c_void_p),
        ("trunk_in", C.c_void_p), ("trunk_out", C.c_void_p),
        ("d1", C.c_void_p), ("d1l", Lay), ("d2", C.c_void_p), ("d2l", Lay), ("d3", C.c_void_p), ("d3l", Lay),
        ("pad_lo", C.c_int32), ("pad_hi", C.c_int32), ("reflect", C.c_int32), ("reserved", C.c_int32),
    ]


class GradSrc(C.Structure):
    _fields_ = [("p", C.c_void_p), ("l", Lay), ("pad_lo", C.c_int32), ("pad_hi", C.c_int32),
                ("reflect", C.c_int32), ("reserved", C.c_int32)]


class GradGather(C.Structure):
    _fields_ = [
        ("nsrc", C.c_int32), ("dst_f32", C.c_int32), ("B", C.c_int32), ("H", C.c_int32),
        ("W", C.c_int32), ("C", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
        ("src", GradSrc * 4), ("trunk", C.c_void_p), ("mask", C.c_void_p), ("ml", Lay),
        ("dst", C.c_void_p), ("dl", Lay),
    ]


class BnBwd(C.Structure):
    _fields_ = [
        ("dz", C.c_void_p), ("dz_f32", C.c_int32), ("relu", C.c_int32), ("dropout", C.c_int32),
        ("drop_key", C.c_uint32), ("x", C.c_void_p), ("xl", Lay), ("coef", C.c_void_p), ("save", C.c_void_p),
        ("sums", C.c_void_p), ("k", C.c_void_p), ("dy", C.c_void_p), ("yl", Lay),
        ("nsrc", C.c_int32), ("reserved", C.c_int32), ("src", GradSrc * 2), ("trunk", C.c_void_p),
    ]


class GateBwd(C.Structure):
    _fields_ = [
        ("dout", C.c_void_p), ("c1", C.c_void_p), ("x2o", C.c_void_p), ("x3o", C.c_void_p), ("sl", Lay),
        ("coef", C.c_void_p), ("save", C.c_void_p), ("sums", C.c_void_p), ("k", C.c_void_p),
        ("ex2", GradSrc), ("ex3", GradSrc),
        ("dy1", C.c_void_p), ("dy2", C.c_void_p), ("dy3", C.c_void_p), ("yl", Lay),
    ]


_lib = None


def load(path=None):
    """Load the shared library (once). ``path`` overrides the in-tree CUDA build (tests only)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MMH_LIB_PATH") or LIB_PATH      # MMH_LIB_PATH: A/B builds of the same library
    if not os.path.exists(p):
        raise MmhError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the mmhand_b200 kernels)" % p)
    lib = C.CDLL(p)
    lib.mmh_last_error.restype = C.c_char_p
    lib.mmh_version.restype = C.c_int
    lib.mmh_is_device_build.restype = C.c_int
    lib.mmh_act_bytes.restype = C.c_int
    lib.mmh_set_pdl.argtypes, lib.mmh_set_pdl.restype = [C.c_int32], C.c_int
    lib.mmh_get_pdl.restype = C.c_int
    lib.act_bytes = lib.mmh_act_bytes()      # 2 (bf16) in the product; 4 only in the fp32 host emulation
    _declare(lib)
    if path is None:
        _lib = lib
    return lib


def _declare(lib):
    vp = C.c_void_p
    lib.mmh_conv_plan_create.argtypes = [C.POINTER(ConvDesc), C.POINTER(vp)]
    lib.mmh_conv_plan_destroy.argtypes = [vp]
    lib.mmh_conv_run.argtypes = [vp, vp]
    lib.mmh_conv_run_key.argtypes = [vp, C.c_uint32, vp]
    lib.mmh_wgrad_plan_create.argtypes = [C.POINTER(WgradDesc), C.POINTER(vp)]
    lib.mmh_wgrad_plan_destroy.argtypes = [vp]
    lib.mmh_wgrad_run.argtypes = [vp, vp]
    for name, args in _SIMPLE_SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name in ("mmh_event_record", "mmh_stream_wait_event"):
        getattr(lib, name)._mmh_no_kernel = True       # stream ordering, not kernels: not counted as launches


_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
_SIMPLE_SIGS = {
    "mmh_assemble_nchw": [_vp, _i32, _vp, _i32, _vp, _vp, _vp, C.POINTER(Lay), _i32, _i32, _i32, _vp],
    "mmh_bn_stats": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mmh_bn_finalize": [_vp, _f32, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _i32, _vp, _vp, _vp],
    "mmh_norm_act": [C.POINTER(NormAct), _vp],
    "mmh_gate_fwd": [C.POINTER(GateFwd), _vp],
    "mmh_grad_gather": [C.POINTER(GradGather), _vp],
    "mmh_bn_bwd_reduce": [C.POINTER(BnBwd), _vp],
    "mmh_bn_bwd_apply": [C.POINTER(BnBwd), _vp],
    "mmh_bn_bwd_finalize": [_vp, _vp, _f32, _vp, _vp, _vp, _i32, _vp],
    "mmh_gate_bwd_reduce": [C.POINTER(GateBwd), _vp],
    "mmh_gate_bwd_apply": [C.POINTER(GateBwd), _vp],
    "mmh_bce_logits": [_vp, _i64, _f32, _f32, _f32, _vp, _vp, _vp],
    "mmh_l1_f32": [_vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp],
    "mmh_perc_loss": [_vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp, _vp],
    "mmh_tanh_bwd": [_vp, _vp, _vp, C.POINTER(Lay), _i32, _vp],
    "mmh_input_grad_nchw": [C.POINTER(GradSrc), _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "mmh_grid_to_nchw": [_vp, C.POINTER(Lay), _vp, _i32, _vp],
    "mmh_pack_weight": [_vp, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _vp],
    "mmh_pack_weight_folded": [_vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp],
    "mmh_unpack_wgrad": [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp],
    "mmh_param_jobs": [_vp, _i32, _i32, _vp],
    "mmh_adam": [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _i32, _f32, _vp],
    "mmh_memset": [_vp, _i32, _i64, _vp],
    "mmh_heatmap_rasterize": [_vp, _i64, _i32, _i32, _f64, _f64, _vp, _vp],
    "mmh_pose_map_rasterize": [_vp, _i64, _i32, _i32, _i32, _f64, _f64, _vp, _vp],
    "mmh_jointsmap_rasterize": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp],
    "mmh_image_pack_bgr8": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mmh_ssim": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp],
    "mmh_image_unpack_u8": [_vp, _i64, _i32, _i32, _i32, _vp, _vp],
    "mmh_depth_unpack_u8": [_vp, _i64, _i32, _i32, _i32, _i32, _f64, _vp, _vp],
    "mmh_peer_create": [_i32, _i32, C.POINTER(_vp)],
    "mmh_peer_handle": [_vp, _vp],
    "mmh_peer_connect": [_vp, _vp],
    "mmh_peer_status": [_vp],
    "mmh_peer_destroy": [_vp],
    "mmh_peer_sum": [_vp, C.c_uint32, _vp, _i32, _vp],
    "mmh_bn_finalize_sync": [_vp, C.c_uint32, _vp, _f32, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp, _vp],
    "mmh_bn_bwd_finalize_sync": [_vp, C.c_uint32, _vp, _vp, _f32, _vp, _vp, _vp, _i32, _vp],
    "mmh_bn_stats_finalize": [_vp, C.c_uint32, _vp, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _f32, _f32, _vp,
                              _vp, _vp],
    "mmh_bn_finalize_reset": [_vp, C.c_uint32, _vp, _f32, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp, _vp],
    "mmh_bn_bwd_reduce_finalize": [_vp, C.c_uint32, C.POINTER(BnBwd), _vp, _f32, _vp, _vp, _vp],
    "mmh_bn_bwd_finalize_reset": [_vp, C.c_uint32, _vp, _f32, _vp, _vp, _vp, _i32, _vp],
    "mmh_gate_bwd_reduce_finalize": [_vp, C.c_uint32, C.POINTER(GateBwd), _vp, _f32, _vp, _vp, _vp],
    "mmh_event_create": [C.POINTER(_vp)],
    "mmh_event_destroy": [_vp],
    "mmh_event_record": [_vp, _vp],
    "mmh_stream_wait_event": [_vp, _vp],
}
PEER_HANDLE_BYTES = 64
EXPORTS = ["mmh_version", "mmh_last_error", "mmh_is_device_build", "mmh_set_pdl", "mmh_get_pdl", "mmh_conv_plan_create", "mmh_conv_plan_destroy",
           "mmh_conv_run", "mmh_conv_run_key", "mmh_wgrad_plan_create", "mmh_wgrad_plan_destroy", "mmh_wgrad_run"] + list(_SIMPLE_SIGS)


def check(lib, rc):
    if rc != 0:
        raise MmhError(lib.mmh_last_error().decode("utf-8", "replace"))


def set_error_check(lib):
    return lambda rc: check(lib, rc)
