"""Isolated timing of the data gradient of a 3x3 reflect convolution (batch 16, 64 x 64) with and without the fused
BatchNorm-backward epilogue (MmhConvDesc.bs_*), next to the standalone reduction kernel it replaces.
  python tools/exp/bs_bench.py            # CUDA events, 30 iterations each
  ncu --set full --import-source on -k regex:conv2_kernel -c 6 python tools/exp/bs_bench.py --iters 2"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mmhand_b200 import convops, lib as L, runtime  # noqa: E402
from mmhand_b200.layouts import geom_s1  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--only", default="")
a = ap.parse_args()
lib = L.load()
ops = runtime.get_ops(torch.device("cuda", 0))
st = lambda: torch.cuda.current_stream().cuda_stream
B, H, W = 16, 64, 64
for Cin, Cout in ((512, 256), (256, 256)):
    g = geom_s1(B, H, W, 3, 'reflect', Cin, Cout)            # the consumer conv c2: Cin -> Cout
    gp = geom_s1(B, H, W, 3, 'reflect', Cin, Cin)            # the producer conv c1 (its raw output feeds BN)
    dy = torch.randn(g.out_lay.rows, Cout, device="cuda").to(torch.bfloat16)
    wd = (torch.randn(9, Cin, Cout, device="cuda") * 0.05).to(torch.bfloat16)
    dx = torch.zeros(g.in_lay.rows, Cin, dtype=torch.bfloat16, device="cuda")
    x = torch.randn(gp.out_lay.rows, Cin, device="cuda").to(torch.bfloat16)
    coef = torch.rand(2 * Cin, device="cuda")
    save = torch.rand(2 * Cin, device="cuda") + 0.5
    sums = torch.zeros(2 * Cin, device="cuda")
    plain = convops.dgrad_plans(lib, g, dy, wd, dx, Cin, Cout)
    fused = convops.dgrad_plans(lib, g, dy, wd, dx, Cin, Cout,
                                bn_bwd=dict(x=x, xl=gp.out_lay, coef=coef, save=save, sums=sums, C=Cin, relu=True,
                                            dropout=True))
    k = torch.zeros(2 * Cin, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    src = [convops_src for convops_src in ()]
    from mmhand_b200.kernels import GradSource

    def reduce():
        ops.bn_bwd_reduce_finalize(None, ([GradSource(dx, g.in_lay, 1, 1, True)], None), False, True, True, 0x1234, x,
                                   gp.out_lay, coef, save, sums, k, ticket, float(B * H * W), None, None)

    runs = (("dgrad plain", lambda: [p.run(st()) for p in plain]),
            ("dgrad fused", lambda: [L.check(lib, lib.mmh_conv_run_key(p.handle, 0x1234, st())) for p in fused]),
            ("bn_bwd_reduce_finalize (replaced)", reduce))
    for name, fn in runs:
        if a.only and a.only not in name:
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        fl = 2.0 * B * H * W * Cin * Cout * 9
        print("%4d->%-4d %-36s %8.1f us  %s" % (Cin, Cout, name, ms * 1000.0,
                                              "%.0f TFLOP/s" % (fl / ms / 1e9) if "dgrad" in name else ""), flush=True)
