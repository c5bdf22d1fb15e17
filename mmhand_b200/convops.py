"""Builders of conv / wgrad plans (C-ABI descriptors) from a ConvGeom and raw buffers.

A *buffer* here is anything with ``data_ptr()`` (a torch tensor on the device of the loaded library).
"""
import ctypes as C

from . import lib as L
from .layouts import ConvGeom, Lay


def _ptr(t, row_off=0, ld=0, elem=2):
    return C.c_void_p(t.data_ptr() + row_off * ld * elem)


class ConvPlan:
    """Owns one MmhConvPlan handle."""

    def __init__(self, lib, desc):
        self.lib = lib
        self.handle = C.c_void_p()
        self.desc = desc
        L.check(lib, lib.mmh_conv_plan_create(C.byref(desc), C.byref(self.handle)))

    def run(self, stream):
        L.check(self.lib, self.lib.mmh_conv_run(self.handle, C.c_void_p(stream)))

    def __del__(self):
        try:
            if self.handle:
                self.lib.mmh_conv_plan_destroy(self.handle)
        except Exception:
            pass


class WgradPlan:
    def __init__(self, lib, desc):
        self.lib = lib
        self.handle = C.c_void_p()
        self.desc = desc
        L.check(lib, lib.mmh_wgrad_plan_create(C.byref(desc), C.byref(self.handle)))

    def run(self, stream):
        L.check(self.lib, self.lib.mmh_wgrad_run(self.handle, C.c_void_p(stream)))

    def __del__(self):
        try:
            if self.handle:
                self.lib.mmh_wgrad_plan_destroy(self.handle)
        except Exception:
            pass


def conv_desc(a, a_rows, a_ld, Cc, w, w_taps, N, taps, M, Hg, Wg, Hv, Wv, out, out_ld, out_row_off=0,
              out_f32=False, zero_invalid=True, bias=None, act=0, out_map=None, n_store=0, bn_sums=None, bn_C=0):
    """taps: list of (weight slot, row shift). out_map: None (identity on the GEMM grid) or a tuple
    (out_img_rows, out_wg, sh, sw, h0, w0)."""
    d = L.ConvDesc()
    d.a = a.data_ptr()
    d.a_rows = a_rows
    d.a_ld = a_ld
    d.C = Cc
    d.w = w.data_ptr()
    d.T = len(taps)
    d.N = N
    d.w_taps = w_taps
    for i, (slot, sh) in enumerate(taps):
        d.w_slot[i] = slot
        d.shift[i] = sh
    d.M = M
    d.Hg, d.Wg, d.Hv, d.Wv = Hg, Wg, Hv, Wv
    esz = out.element_size()
    d.out = out.data_ptr() + out_row_off * out_ld * esz
    d.out_f32 = 1 if out_f32 else 0
    d.out_ld = out_ld
    if out_map is None:
        d.out_img_rows, d.out_wg, d.out_sh, d.out_sw, d.out_h0, d.out_w0 = Hg * Wg, Wg, 1, 1, 0, 0
    else:
        d.out_img_rows, d.out_wg, d.out_sh, d.out_sw, d.out_h0, d.out_w0 = out_map
    d.zero_invalid = 1 if zero_invalid else 0
    d.bias = bias.data_ptr() if bias is not None else None
    d.act = act
    d.n_store = n_store
    d.bn_sums = bn_sums.data_ptr() if bn_sums is not None else None
    d.bn_C = bn_C if bn_sums is not None else 0
    return d


def fwd_plans(lib, g: ConvGeom, a_buf, w_packed, out_buf, Cin_p, Cout_p, bias=None, act=0, out_f32=False,
              out_map=None, zero_invalid=True, bn_sums=None, bn_C=0):
    """Plans of the forward convolution described by ``g`` (one, or four for a transposed conv)."""
    plans = []
    il, ol = g.in_lay, g.out_lay
    M = il.plane_rows
    for ln in g.fwd:
        if g.kind == 'up':
            Hv, Wv = g.H, g.W
        else:
            Hv, Wv = g.Ho, g.Wo
        d = conv_desc(a_buf, il.rows, il.ld, Cin_p, w_packed, g.k * g.k, Cout_p, ln.taps, M, il.Hg, il.Wg, Hv, Wv,
                      out_buf, ol.ld, out_row_off=ln.out_plane * ol.plane_rows if g.kind == 'up' else 0,
                      out_f32=out_f32, zero_invalid=zero_invalid, bias=bias, act=act, out_map=out_map,
                      bn_sums=bn_sums, bn_C=bn_C)
        plans.append(ConvPlan(lib, d))
    return plans


def dgrad_plans(lib, g: ConvGeom, dy_buf, wd_packed, dx_buf, Cin_p, Cout_p, dx_ld=None, bn_bwd=None):
    """Plans of the data gradient: A = dY (layout g.out_lay), output = gradient of g.in_lay (every row).

    bn_bwd: optional dict(x, xl, coef, save, sums, C, relu, dropout) -- the convolution's input came from
    conv_p -> BatchNorm -> [ReLU] -> [Dropout] -> reflect padding: the launch stores the masked gradient and accumulates the
    BatchNorm-backward sums of that producer (MmhConvDesc.bs_*; stride-1 reflect geometries only)."""
    plans = []
    il, ol = g.in_lay, g.out_lay
    M = il.plane_rows
    ld = dx_ld or il.ld
    for ln in g.bwd:
        d = conv_desc(dy_buf, ol.rows, ol.ld, Cout_p, wd_packed, g.k * g.k, Cin_p, ln.taps, M, il.Hg, il.Wg, il.Hg,
                      il.Wg, dx_buf, ld, out_row_off=ln.out_plane * il.plane_rows if g.kind == 's2' else 0,
                      zero_invalid=False)
        if bn_bwd is not None:
            assert g.kind == 's1' and g.pad_mode == 'reflect' and len(g.bwd) == 1
            xl = bn_bwd["xl"]
            assert not xl.phase and xl.h0 == 0 and xl.w0 == 0 and xl.c0 == 0 and (xl.H, xl.W) == (il.H, il.W)
            d.bs_x = bn_bwd["x"].data_ptr()
            d.bs_coef, d.bs_save = bn_bwd["coef"].data_ptr(), bn_bwd["save"].data_ptr()
            d.bs_sums = bn_bwd["sums"].data_ptr()
            d.bs_x_ld, d.bs_xHg, d.bs_xWg = xl.ld, xl.Hg, xl.Wg
            d.bs_H, d.bs_W, d.bs_pad, d.bs_C = il.H, il.W, g.pad, bn_bwd["C"]
            d.bs_relu, d.bs_dropout, d.bs_drop_key = int(bn_bwd["relu"]), int(bn_bwd["dropout"]), 0
        plans.append(ConvPlan(lib, d))
    return plans


def wgrad_plans(lib, g: ConvGeom, a_buf, dy_buf, dw_buf, Cin_p, Cout_p, Cin, Cout, split_k=0):
    """dw_buf: fp32 [k*k][Cout][Cin] (packed-gradient layout, caller zeroes)."""
    plans = []
    il, ol = g.in_lay, g.out_lay
    M = il.plane_rows
    for ln in g.fwd:
        d = L.WgradDesc()
        d.a = a_buf.data_ptr()
        d.a_rows = il.rows
        d.a_ld = il.ld
        d.C = Cin_p
        off = ln.out_plane * ol.plane_rows if g.kind == 'up' else 0
        d.dy = dy_buf.data_ptr() + off * ol.ld * dy_buf.element_size()
        d.M = M
        d.dy_ld = ol.ld
        d.N = Cout_p
        d.T = len(ln.taps)
        for i, (slot, sh) in enumerate(ln.taps):
            d.shift[i] = sh
            d.tap_index[i] = slot
        d.dw = dw_buf.data_ptr()
        d.dw_taps = g.k * g.k
        d.N_store = Cout
        d.C_store = Cin
        d.split_k = split_k
        plans.append(WgradPlan(lib, d))
    return plans
