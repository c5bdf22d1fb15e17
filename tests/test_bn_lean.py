"""Lean BatchNorm-backward kernels (csrc/elementwise.cu: BnLeanReduceF / BnLeanApplyF, row launchers of ew_framework.h)
against the general functors: same sums and the same data gradient for plain, reflect-folded and parity-plane sources,
with and without ReLU / dropout masks -- on the host emulation here, on the CUDA library under -m gpu."""
import os

import pytest
import torch

from mmhand_b200.kernels import GradSource
from mmhand_b200.layouts import geom_s1, geom_s2, geom_up
import hostemu


def _case(ops, kind, relu, drop, dev="cpu", B=2, H=8, W=12, Cc=16):
    g = torch.Generator().manual_seed(5)
    dt = torch.float32 if ops.lib.act_bytes == 4 else torch.bfloat16
    if kind == "reflect":          # producer 3x3 conv -> BN -> consumer 3x3 reflect conv
        gp = geom_s1(B, H, W, 3, 'reflect', Cc, Cc)
        gc = geom_s1(B, H, W, 3, 'reflect', Cc, Cc)
        xl, sl, lo, hi, refl = gp.out_lay, gc.in_lay, 1, 1, True
    elif kind == "s2":             # consumer is a stride-2 conv: its input gradient lives in four parity planes
        gp = geom_s1(B, H, W, 7, 'reflect', Cc, Cc)
        gc = geom_s2(B, H, W, Cc, Cc)
        xl, sl, lo, hi, refl = gp.out_lay, gc.in_lay, 1, 1, False
    else:                          # producer is a transposed conv (raw output in parity planes), consumer 7x7 reflect
        gp = geom_up(B, H // 2, W // 2, Cc, Cc)
        gc = geom_s1(B, H, W, 7, 'reflect', Cc, Cc)
        xl, sl, lo, hi, refl = gp.out_lay, gc.in_lay, 3, 3, True
    mk = lambda rows, ld: (torch.randn(rows, ld, generator=g)).to(dt).to(dev)
    x, src = mk(xl.rows, xl.ld), mk(sl.rows, sl.ld)
    coef = torch.cat([torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.3]).to(dev)
    save = torch.cat([torch.randn(Cc, generator=g) * 0.2, torch.rand(Cc, generator=g) + 0.5]).to(dev)
    k = (torch.randn(2 * Cc, generator=g) * 0.1).to(dev)
    out = {}
    for lean in ("0", "1"):
        os.environ["MMH_BN_LEAN"] = lean
        sums = torch.zeros(2 * Cc, device=dev)
        dy = torch.zeros(xl.rows, xl.ld, dtype=dt, device=dev)
        dz = ([GradSource(src, sl, lo, hi, refl)], None)
        ops.bn_bwd_reduce(dz, False, relu, drop, 0x51F3, x, xl, coef, save, sums)
        ops.bn_bwd_apply(dz, False, relu, drop, 0x51F3, x, xl, coef, save, k, dy, xl)
        if dev != "cpu":
            torch.cuda.synchronize()
        out[lean] = (sums.float().cpu(), dy.float().cpu())
    os.environ.pop("MMH_BN_LEAN", None)
    return out


@pytest.mark.parametrize("kind", ["reflect", "s2", "up"])
@pytest.mark.parametrize("relu,drop", [(True, True), (True, False), (False, False)])
def test_lean_matches_general_hostemu(kind, relu, drop):
    ops = hostemu.ops(f32=True)
    o = _case(ops, kind, relu, drop)
    (s0, d0), (s1, d1) = o["0"], o["1"]
    assert torch.isfinite(s0).all() and s0.abs().max() > 0
    assert torch.allclose(s0, s1, rtol=2e-5, atol=2e-4), (s0 - s1).abs().max()
    assert torch.equal(d0, d1) or torch.allclose(d0, d1, rtol=1e-6, atol=1e-6)



@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["reflect", "s2", "up"])
def test_lean_matches_general_gpu(kind):
    from mmhand_b200 import runtime
    ops = runtime.get_ops(torch.device("cuda", 0))
    o = _case(ops, kind, True, True, dev="cuda", B=2, H=16, W=24, Cc=64)
    (s0, d0), (s1, d1) = o["0"], o["1"]
    assert s0.abs().max() > 0
    assert torch.allclose(s0, s1, rtol=1e-4, atol=1e-3 * s0.abs().max().item()), (s0 - s1).abs().max()
    assert (d0 - d1).abs().max() <= 1e-2 * d0.abs().max()
