"""GPU parity cases for the tcgen05 conv / wgrad kernels against torch fp32 convolutions on the same
bf16-rounded operands. Used by tests/test_gpu_conv.py (pytest -m gpu) and tools/bringup.py (one subprocess
per case, so one faulting kernel cannot take the others down).

torch here is the *oracle* (the reference's arithmetic is torch.nn.Conv2d / ConvTranspose2d,
models/Generator.py:62-111,158-259); the code under test is libmmhand_sm100.so through the C ABI.
"""
import torch
import torch.nn.functional as F

from mmhand_b200 import convops, lib as L
from mmhand_b200.layouts import chan_pad, geom_s1, geom_s2, geom_up

import os

# MMH_TEST_HOSTEMU=1: run the same cases against the host emulation (validates the harness, the tap
# tables and the layouts on a CPU-only box); default: the CUDA library on cuda:0.
HOSTEMU = os.environ.get("MMH_TEST_HOSTEMU", "0") == "1"
DEV = "cpu" if HOSTEMU else "cuda"


def _lib():
    if HOSTEMU:
        import hostemu
        return hostemu.load()
    return L.load()


def _stream():
    return 0 if HOSTEMU else torch.cuda.current_stream().cuda_stream


def _sync():
    if not HOSTEMU:
        torch.cuda.synchronize()


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).float()


def pack_w(w, Np, Cp, transposed=False, swap=False):
    """OIHW fp32 (or IOHW for ConvTranspose2d when transposed=True) -> bf16 [k*k][Np][Cp].
    swap=True packs the data-gradient operand (roles of the two channel dims exchanged)."""
    if transposed:
        w = w.permute(1, 0, 2, 3)           # -> [co][ci][kh][kw]
    if swap:
        w = w.permute(1, 0, 2, 3)
    n, c, kh, kw = w.shape
    out = torch.zeros(kh * kw, Np, Cp, dtype=torch.float32)
    out[:, :n, :c] = w.permute(2, 3, 0, 1).reshape(kh * kw, n, c)
    return out.to(torch.bfloat16).to(DEV).contiguous()


def to_grid(x_nchw, lay, Cp, pad_lo, pad_hi, mode):
    """Logical NCHW fp32 -> grid buffer [rows, ld] bf16 following ``lay`` (halo by ``mode``)."""
    B, C, H, W = x_nchw.shape
    if mode == 'reflect':
        xp = F.pad(x_nchw, (pad_lo, pad_hi, pad_lo, pad_hi), mode='reflect')
    else:
        xp = F.pad(x_nchw, (pad_lo, pad_hi, pad_lo, pad_hi))
    xp = xp.permute(0, 2, 3, 1)             # B, Hp, Wp, C
    Hp, Wp = xp.shape[1], xp.shape[2]
    buf = torch.zeros(lay.rows, lay.ld, dtype=torch.float32)
    if not lay.phase:
        full = torch.zeros(B, lay.Hg, lay.Wg, lay.ld)
        h0, w0 = lay.h0 - pad_lo, lay.w0 - pad_lo
        full[:, h0:h0 + Hp, w0:w0 + Wp, :C] = xp
        buf = full.reshape(lay.rows, lay.ld)
    else:
        assert lay.h0 == pad_lo and lay.w0 == pad_lo
        planes = torch.zeros(4, B, lay.Hg, lay.Wg, lay.ld)
        for ph in range(2):
            for pw in range(2):
                sub = xp[:, ph::2, pw::2]
                planes[ph * 2 + pw, :, :sub.shape[1], :sub.shape[2], :C] = sub
        buf = planes.reshape(lay.rows, lay.ld)
    return buf.to(torch.bfloat16).to(DEV).contiguous()


def from_grid(buf, lay, C, H, W, h_lo=0, w_lo=0):
    """Grid buffer -> logical NCHW fp32 for the window [h_lo, h_lo+H) x [w_lo, w_lo+W) of logical coords."""
    buf = buf.float().cpu()
    B = lay.B
    if not lay.phase:
        full = buf.reshape(B, lay.Hg, lay.Wg, lay.ld)
        h0, w0 = lay.h0 + h_lo, lay.w0 + w_lo
        return full[:, h0:h0 + H, w0:w0 + W, :C].permute(0, 3, 1, 2).contiguous()
    planes = buf.reshape(4, B, lay.Hg, lay.Wg, lay.ld)
    out = torch.zeros(B, H, W, lay.ld)
    for h in range(H):
        hp = h + h_lo + lay.h0
        for w in range(W):
            wp = w + w_lo + lay.w0
            out[:, h, w] = planes[(hp & 1) * 2 + (wp & 1), :, hp >> 1, wp >> 1]
    return out[..., :C].permute(0, 3, 1, 2).contiguous()


def _cmp(name, got, want, tol=2e-2):
    scale = want.abs().max().item() + 1e-6
    err = (got - want).abs().max().item()
    ok = bool(err <= tol * scale) and bool(torch.isfinite(got).all())
    return {"case": name, "ok": ok, "max_err": err, "ref_scale": scale}


# -------------------------------------------------------------------------------------------------
def case_gemm(M=1024, Cc=64, N=64, seed=1):
    """T=1, shift 0: plain [M,C] x [N,C]^T GEMM."""
    lib = _lib()
    a = _rand(M, Cc, seed=seed)
    w = _rand(N, Cc, scale=0.1, seed=seed + 1)
    a_d, w_d = a.to(torch.bfloat16).to(DEV), w.to(torch.bfloat16).to(DEV)
    out = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    d = convops.conv_desc(a_d, M, Cc, Cc, w_d, 1, N, [(0, 0)], M, 1, M, 1, M, out, N)
    plan = convops.ConvPlan(lib, d)
    plan.run(_stream())
    _sync()
    return _cmp("gemm_M%d_C%d_N%d" % (M, Cc, N), out.float().cpu(), a @ w.t())


def _conv_setup(kind, B, H, W, Cin, Cout, k, mode, seed):
    Cin_p, Cout_p = chan_pad(Cin), chan_pad(Cout)
    if kind == 's1':
        g = geom_s1(B, H, W, k, mode, Cin_p, Cout_p)
    elif kind == 's2':
        g = geom_s2(B, H, W, Cin_p, Cout_p)
    else:
        g = geom_up(B, H, W, Cin_p, Cout_p)
    x = _rand(B, Cin, H, W, seed=seed)
    if kind == 'up':
        w = _rand(Cin, Cout, 3, 3, scale=0.1, seed=seed + 1)
    else:
        w = _rand(Cout, Cin, k, k, scale=0.1, seed=seed + 1)
    return g, x, w, Cin_p, Cout_p


def _ref_dev(*ts):
    """Big cases: the torch fp32 reference runs on the GPU too (TF32 off), results come back to the host."""
    if HOSTEMU or not BIG_ON_GPU or ts[0].numel() < (1 << 22):
        return ts
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return tuple(t.to(DEV) for t in ts)


BIG_ON_GPU = True


def _ref_fwd(kind, x, w, k, mode):
    p = (k - 1) // 2
    if kind == 's1':
        xp = F.pad(x, (p, p, p, p), mode='reflect') if mode == 'reflect' else F.pad(x, (p, p, p, p))
        return F.conv2d(xp, w)
    if kind == 's2':
        return F.conv2d(x, w, stride=2, padding=1)
    return F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)


def case_conv_fwd(kind='s1', B=2, H=16, W=16, Cin=64, Cout=64, k=3, mode='reflect', seed=3, bias=False, act=0):
    lib = _lib()
    g, x, w, Cin_p, Cout_p = _conv_setup(kind, B, H, W, Cin, Cout, k, mode, seed)
    a_buf = to_grid(x, g.in_lay, Cin_p, g.in_pad_lo, g.in_pad_hi, mode if kind == 's1' else 'zero')
    wp = pack_w(w, Cout_p, Cin_p, transposed=(kind == 'up'))
    out = torch.full((g.out_lay.rows, Cout_p), 7.0, dtype=torch.bfloat16, device=DEV)
    b = None
    bt = None
    if bias:
        bt = _rand(Cout, seed=seed + 5)
        b = torch.zeros(Cout_p)
        b[:Cout] = bt
        b = b.to(DEV)
    plans = convops.fwd_plans(lib, g, a_buf, wp, out, Cin_p, Cout_p, bias=b, act=act)
    for p_ in plans:
        p_.run(_stream())
    _sync()
    want = _ref_fwd(kind, *_ref_dev(x, w), k, mode).cpu()
    if bt is not None:
        want = want + bt.view(1, -1, 1, 1)
    if act == 1:
        want = want.relu()
    elif act == 2:
        want = want.tanh()
    got = from_grid(out, g.out_lay, Cout, g.Ho, g.Wo)
    r = _cmp("fwd_%s_B%d_%dx%d_%d-%d_k%d_%s" % (kind, B, H, W, Cin, Cout, k, mode), got, want)
    # zero_invalid contract: every grid position outside the valid window is exactly 0
    ol = g.out_lay
    full = out.float().cpu().reshape(4 if ol.phase else 1, B, ol.Hg, ol.Wg, Cout_p).clone()
    hv, wv = (g.H, g.W) if kind == 'up' else (g.Ho, g.Wo)
    full[:, :, :hv, :wv, :] = 0
    r["invalid_abs_max"] = full.abs().max().item()
    r["ok"] = r["ok"] and r["invalid_abs_max"] == 0.0
    return r


def case_conv_dgrad(kind='s1', B=2, H=16, W=16, Cin=64, Cout=64, k=3, mode='reflect', seed=5):
    lib = _lib()
    g, x, w, Cin_p, Cout_p = _conv_setup(kind, B, H, W, Cin, Cout, k, mode, seed)
    dy = _rand(B, Cout, g.Ho, g.Wo, seed=seed + 2)
    # reference: gradient w.r.t. the *padded* input (the conv's A operand)
    p = (k - 1) // 2
    xr, wr, dyr = _ref_dev(x, w, dy)
    if kind == 's1':
        xp = F.pad(xr, (p, p, p, p)).requires_grad_(True)
        y = F.conv2d(xp, wr)
    elif kind == 's2':
        xp = F.pad(xr, (1, 1, 1, 1)).requires_grad_(True)
        y = F.conv2d(xp, wr, stride=2)
    else:
        xp = xr.clone().requires_grad_(True)
        y = F.conv_transpose2d(xp, wr, stride=2, padding=1, output_padding=1)
    (want,) = torch.autograd.grad(y, xp, dyr)
    want = want.cpu()
    dy_buf = to_grid(dy, g.out_lay, Cout_p, 0, 0, 'zero')
    wd = pack_w(w, Cin_p, Cout_p, transposed=(kind == 'up'), swap=True)
    dx = torch.full((g.in_lay.rows, Cin_p), 7.0, dtype=torch.bfloat16, device=DEV)
    plans = convops.dgrad_plans(lib, g, dy_buf, wd, dx, Cin_p, Cout_p)
    for p_ in plans:
        p_.run(_stream())
    _sync()
    if kind == 'up':
        got = from_grid(dx, g.in_lay, Cin, H, W)
    else:
        got = from_grid(dx, g.in_lay, Cin, H + 2 * p, W + 2 * p, -p, -p)
    return _cmp("dgrad_%s_B%d_%dx%d_%d-%d_k%d" % (kind, B, H, W, Cin, Cout, k), got, want)


def case_conv_dgrad_bnbwd(B=2, H=16, W=16, Cin=64, Cout=64, relu=True, dropout=True, seed=11, key=0x1234ABCD):
    """Data gradient of a 3x3 reflect convolution with the fused BatchNorm-backward epilogue (MmhConvDesc.bs_*): the
    stored rows are the masked gradient of the mirrored pixel and bs_sums holds (sum dze, sum dze * xhat)."""
    from oracle.patn_ref import dropout_mask
    lib = _lib()
    k, mode = 3, 'reflect'
    g, _, w, Cin_p, Cout_p = _conv_setup('s1', B, H, W, Cin, Cout, k, mode, seed)
    dy = _rand(B, Cout, H, W, seed=seed + 2)
    xraw = _rand(B, Cin, H, W, seed=seed + 3)                 # raw output of the producer convolution
    gen = torch.Generator().manual_seed(seed + 4)
    a = torch.rand(Cin, generator=gen) + 0.5
    b = torch.randn(Cin, generator=gen) * 0.5
    mean = torch.randn(Cin, generator=gen) * 0.3
    rstd = torch.rand(Cin, generator=gen) + 0.5
    xp = F.pad(torch.zeros(B, Cin, H, W), (1, 1, 1, 1)).requires_grad_(True)
    (dxp,) = torch.autograd.grad(F.conv2d(xp, w), xp, dy)    # gradient w.r.t. the padded input
    mask = torch.ones(B, Cin, H, W)
    if relu:
        mask = mask * ((a.view(1, -1, 1, 1) * xraw + b.view(1, -1, 1, 1)) > 0).float()
    if dropout:
        mask = mask * 2.0 * dropout_mask((B, Cin, H, W), key)
    rp = lambda t: F.pad(t, (1, 1, 1, 1), mode='reflect')
    want = (dxp * rp(mask)).to(torch.bfloat16).float()
    xhat = (xraw - mean.view(1, -1, 1, 1)) * rstd.view(1, -1, 1, 1)
    want_s0 = want.sum(dim=(0, 2, 3))
    want_s1 = (want * rp(xhat)).sum(dim=(0, 2, 3))
    # producer's raw output on ITS output grid (same Hg x Wg, content at (0, 0)), statistics / coefficient vectors
    gp = geom_s1(B, H, W, 3, 'reflect', Cin_p, Cin_p)
    x_buf = to_grid(xraw, gp.out_lay, Cin_p, 0, 0, 'zero')
    coef = torch.zeros(2 * Cin); coef[:Cin] = a; coef[Cin:] = b
    save = torch.zeros(2 * Cin); save[:Cin] = mean; save[Cin:] = rstd
    coef, save = coef.to(DEV), save.to(DEV)
    sums = torch.zeros(2 * Cin, dtype=torch.float32, device=DEV)
    dy_buf = to_grid(dy, g.out_lay, Cout_p, 0, 0, 'zero')
    wd = pack_w(w, Cin_p, Cout_p, swap=True)
    dx = torch.full((g.in_lay.rows, Cin_p), 7.0, dtype=torch.bfloat16, device=DEV)
    plans = convops.dgrad_plans(lib, g, dy_buf, wd, dx, Cin_p, Cout_p,
                                bn_bwd=dict(x=x_buf, xl=gp.out_lay, coef=coef, save=save, sums=sums, C=Cin, relu=relu,
                                            dropout=dropout))
    for p_ in plans:
        L.check(lib, lib.mmh_conv_run_key(p_.handle, key, _stream()))
    _sync()
    got = from_grid(dx, g.in_lay, Cin, H + 2, W + 2, -1, -1)
    r = _cmp("dgrad_bnbwd_B%d_%dx%d_%d-%d_r%d_d%d" % (B, H, W, Cin, Cout, relu, dropout), got, want)
    s = sums.float().cpu()
    # the masks must be decided identically (a flipped mask shows as an O(1) relative error on that element)
    e0 = (s[:Cin] - want_s0).abs().max().item() / (want.abs().sum(dim=(0, 2, 3)).max().item() + 1e-6)
    e1 = (s[Cin:] - want_s1).abs().max().item() / ((want * rp(xhat)).abs().sum(dim=(0, 2, 3)).max().item() + 1e-6)
    r["sum_err"], r["sumx_err"] = e0, e1
    r["ok"] = r["ok"] and e0 <= 2e-3 and e1 <= 2e-3
    return r


def case_conv_wgrad(kind='s1', B=2, H=16, W=16, Cin=64, Cout=64, k=3, mode='reflect', seed=7, split_k=0):
    lib = _lib()
    g, x, w, Cin_p, Cout_p = _conv_setup(kind, B, H, W, Cin, Cout, k, mode, seed)
    dy = _rand(B, Cout, g.Ho, g.Wo, seed=seed + 2)
    xr, wr, dyr = _ref_dev(x, w, dy)
    wv = wr.clone().requires_grad_(True)
    y = _ref_fwd(kind, xr, wv, k, mode)
    (want,) = torch.autograd.grad(y, wv, dyr)
    want = want.cpu()
    a_buf = to_grid(x, g.in_lay, Cin_p, g.in_pad_lo, g.in_pad_hi, mode if kind == 's1' else 'zero')
    dy_buf = to_grid(dy, g.out_lay, Cout_p, 0, 0, 'zero')
    dw = torch.zeros(k * k, Cout, Cin, dtype=torch.float32, device=DEV)
    plans = convops.wgrad_plans(lib, g, a_buf, dy_buf, dw, Cin_p, Cout_p, Cin, Cout, split_k=split_k)
    for p_ in plans:
        p_.run(_stream())
    _sync()
    got = dw.cpu().reshape(k, k, Cout, Cin).permute(2, 3, 0, 1)
    if kind == 'up':
        got = got.permute(1, 0, 2, 3)
    r = _cmp("wgrad_%s_B%d_%dx%d_%d-%d_k%d" % (kind, B, H, W, Cin, Cout, k), got.contiguous(), want,
             tol=1e-2)
    return r


def case_perf(B=16, H=64, W=64, Cin=256, Cout=256, iters=20, k=3):
    """Device time of the dominant 3x3 layer (fprop, dgrad, wgrad). MMH_PERF_ITERS overrides iters (ncu runs)."""
    iters = int(os.environ.get("MMH_PERF_ITERS", iters))
    warm = 1 if iters == 1 else 3
    lib = _lib()
    Cin_p, Cout_p = chan_pad(Cin), chan_pad(Cout)
    g = geom_s1(B, H, W, k, 'reflect', Cin_p, Cout_p)
    a_buf = torch.randn(g.in_lay.rows, Cin_p, device=DEV).to(torch.bfloat16)
    wp = (torch.randn(k * k, Cout_p, Cin_p, device=DEV) * 0.05).to(torch.bfloat16)
    wd = (torch.randn(k * k, Cin_p, Cout_p, device=DEV) * 0.05).to(torch.bfloat16)
    out = torch.zeros(g.out_lay.rows, Cout_p, dtype=torch.bfloat16, device=DEV)
    dx = torch.zeros(g.in_lay.rows, Cin_p, dtype=torch.bfloat16, device=DEV)
    dw = torch.zeros(k * k, Cout, Cin, dtype=torch.float32, device=DEV)
    res = {"case": "perf_B%d_%dx%d_%d-%d_k%d" % (B, H, W, Cin, Cout, k), "ok": True}
    flops = 2.0 * B * H * W * Cin_p * Cout_p * k * k
    for name, plans in (("fprop", convops.fwd_plans(lib, g, a_buf, wp, out, Cin_p, Cout_p)),
                        ("dgrad", convops.dgrad_plans(lib, g, out, wd, dx, Cin_p, Cout_p)),
                        ("wgrad", convops.wgrad_plans(lib, g, a_buf, out, dw, Cin_p, Cout_p, Cin, Cout))):
        for _ in range(warm):
            for p_ in plans:
                p_.run(_stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _sync()
        e0.record()
        for _ in range(iters):
            for p_ in plans:
                p_.run(_stream())
        e1.record()
        _sync()
        ms = e0.elapsed_time(e1) / iters
        res[name + "_ms"] = ms
        res[name + "_tflops"] = flops / ms / 1e9
    return res


CASES = {
    "gemm_64": lambda: case_gemm(1024, 64, 64),
    "gemm_256": lambda: case_gemm(4096, 256, 256),
    "gemm_512": lambda: case_gemm(2048, 512, 512),
    "gemm_c16": lambda: case_gemm(1000, 16, 16),
    "gemm_c32": lambda: case_gemm(1000, 32, 32),
    "gemm_c48": lambda: case_gemm(1000, 48, 64),
    "fwd_s1_3x3": lambda: case_conv_fwd('s1', 2, 16, 16, 64, 64, 3, 'reflect'),
    "fwd_s1_3x3_256": lambda: case_conv_fwd('s1', 2, 64, 64, 256, 256, 3, 'reflect'),
    "fwd_s1_3x3_512": lambda: case_conv_fwd('s1', 1, 64, 64, 512, 256, 3, 'reflect'),
    "fwd_s1_7x7_c3": lambda: case_conv_fwd('s1', 2, 32, 32, 3, 64, 7, 'reflect'),
    "fwd_s1_7x7_c42": lambda: case_conv_fwd('s1', 1, 32, 32, 42, 64, 7, 'reflect'),
    "fwd_s1_7x7_c24": lambda: case_conv_fwd('s1', 1, 32, 32, 24, 64, 7, 'reflect'),
    "fwd_s1_7x7_out3": lambda: case_conv_fwd('s1', 1, 32, 32, 64, 3, 7, 'reflect', bias=True, act=2),
    "fwd_vgg1": lambda: case_conv_fwd('s1', 1, 32, 32, 3, 64, 3, 'zero', bias=True, act=1),
    "fwd_s2": lambda: case_conv_fwd('s2', 2, 32, 32, 64, 128),
    "fwd_s2_b": lambda: case_conv_fwd('s2', 1, 32, 32, 128, 256),
    "fwd_up": lambda: case_conv_fwd('up', 2, 16, 16, 256, 128),
    "fwd_up_b": lambda: case_conv_fwd('up', 1, 32, 32, 128, 64),
    "dgrad_s1_3x3": lambda: case_conv_dgrad('s1', 2, 16, 16, 64, 64, 3),
    "dgrad_s1_3x3_512": lambda: case_conv_dgrad('s1', 1, 32, 32, 512, 256, 3),
    "dgrad_s1_7x7_c24": lambda: case_conv_dgrad('s1', 1, 32, 32, 24, 64, 7),
    "dgrad_s1_7x7_out3": lambda: case_conv_dgrad('s1', 1, 32, 32, 64, 3, 7),
    "dgrad_s2": lambda: case_conv_dgrad('s2', 2, 32, 32, 64, 128),
    "dgrad_up": lambda: case_conv_dgrad('up', 2, 16, 16, 256, 128),
    "dgrad_bnbwd_64": lambda: case_conv_dgrad_bnbwd(2, 16, 16, 64, 64),
    "dgrad_bnbwd_256": lambda: case_conv_dgrad_bnbwd(2, 64, 64, 256, 256),
    "dgrad_bnbwd_512": lambda: case_conv_dgrad_bnbwd(3, 64, 64, 512, 256),
    "dgrad_bnbwd_256_norelu": lambda: case_conv_dgrad_bnbwd(1, 32, 32, 256, 256, relu=False, dropout=False),
    "dgrad_bnbwd_128_nodrop": lambda: case_conv_dgrad_bnbwd(5, 24, 40, 128, 64, relu=True, dropout=False),
    "wgrad_s1_3x3": lambda: case_conv_wgrad('s1', 2, 16, 16, 64, 64, 3),
    "wgrad_s1_3x3_256": lambda: case_conv_wgrad('s1', 2, 32, 32, 256, 256, 3),
    "wgrad_s1_3x3_512": lambda: case_conv_wgrad('s1', 1, 32, 32, 512, 256, 3),
    "wgrad_s1_7x7_c3": lambda: case_conv_wgrad('s1', 1, 32, 32, 3, 64, 7),
    "wgrad_s1_7x7_c42": lambda: case_conv_wgrad('s1', 1, 32, 32, 42, 64, 7),
    "wgrad_s1_7x7_out3": lambda: case_conv_wgrad('s1', 1, 32, 32, 64, 3, 7),
    "wgrad_s1_c32": lambda: case_conv_wgrad('s1', 1, 32, 32, 24, 64, 7),
    "wgrad_s2": lambda: case_conv_wgrad('s2', 2, 32, 32, 64, 128),
    "wgrad_up": lambda: case_conv_wgrad('up', 2, 16, 16, 256, 128),
    # full-size cases: the batch-16 / 256 x 256 shapes of BASELINE configs[2] (M up to 1.1 M grid rows, deep split-K)
    "big_fwd_stem_c3": lambda: case_conv_fwd('s1', 16, 256, 256, 3, 64, 7, 'reflect'),
    "big_fwd_stem_c42": lambda: case_conv_fwd('s1', 16, 256, 256, 42, 64, 7, 'reflect'),
    "big_fwd_out3": lambda: case_conv_fwd('s1', 16, 256, 256, 64, 3, 7, 'reflect', bias=True, act=2),
    "big_fwd_s2": lambda: case_conv_fwd('s2', 16, 256, 256, 64, 128),
    "big_fwd_s2_b": lambda: case_conv_fwd('s2', 16, 128, 128, 128, 256),
    "big_fwd_3x3_256": lambda: case_conv_fwd('s1', 16, 64, 64, 256, 256, 3, 'reflect'),
    "big_fwd_3x3_512": lambda: case_conv_fwd('s1', 16, 64, 64, 512, 512, 3, 'reflect'),
    "big_fwd_up": lambda: case_conv_fwd('up', 16, 64, 64, 256, 128),
    "big_fwd_up_b": lambda: case_conv_fwd('up', 16, 128, 128, 128, 64),
    "big_dgrad_out3": lambda: case_conv_dgrad('s1', 16, 256, 256, 64, 3, 7),
    "big_dgrad_stem_c24": lambda: case_conv_dgrad('s1', 16, 256, 256, 24, 64, 7),
    "big_dgrad_s2": lambda: case_conv_dgrad('s2', 16, 256, 256, 64, 128),
    "big_dgrad_s2_b": lambda: case_conv_dgrad('s2', 16, 128, 128, 128, 256),
    "big_dgrad_3x3_512": lambda: case_conv_dgrad('s1', 16, 64, 64, 512, 256, 3),
    "big_dgrad_up": lambda: case_conv_dgrad('up', 16, 64, 64, 256, 128),
    "big_dgrad_up_b": lambda: case_conv_dgrad('up', 16, 128, 128, 128, 64),
    "big_dgrad_bnbwd_512": lambda: case_conv_dgrad_bnbwd(16, 64, 64, 512, 256),
    "big_dgrad_bnbwd_256": lambda: case_conv_dgrad_bnbwd(16, 64, 64, 256, 256),
    "big_wgrad_stem_c3": lambda: case_conv_wgrad('s1', 16, 256, 256, 3, 64, 7),
    "big_wgrad_stem_c42": lambda: case_conv_wgrad('s1', 16, 256, 256, 42, 64, 7),
    "big_wgrad_stem_c24": lambda: case_conv_wgrad('s1', 16, 256, 256, 24, 64, 7),
    "big_wgrad_out3": lambda: case_conv_wgrad('s1', 16, 256, 256, 64, 3, 7),
    "big_wgrad_s2": lambda: case_conv_wgrad('s2', 16, 256, 256, 64, 128),
    "big_wgrad_s2_b": lambda: case_conv_wgrad('s2', 16, 128, 128, 128, 256),
    "big_wgrad_3x3_256": lambda: case_conv_wgrad('s1', 16, 64, 64, 256, 256, 3),
    "big_wgrad_3x3_512": lambda: case_conv_wgrad('s1', 16, 64, 64, 512, 512, 3),
    "big_wgrad_up": lambda: case_conv_wgrad('up', 16, 64, 64, 256, 128),
    "big_wgrad_up_b": lambda: case_conv_wgrad('up', 16, 128, 128, 128, 64),
    "perf": lambda: case_perf(),
    "perf512": lambda: case_perf(16, 64, 64, 512, 512),
    "perf_stem": lambda: case_perf(16, 256, 256, 3, 64, k=7),
    "perf_stem42": lambda: case_perf(16, 256, 256, 42, 64, k=7),
    "perf_out": lambda: case_perf(16, 256, 256, 64, 3, k=7),
    "perf_d1": lambda: case_perf(16, 128, 128, 128, 128),
}
