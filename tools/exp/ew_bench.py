"""Isolated timing of the BatchNorm-backward pair (reduce + finalise, apply) and of norm_act on the training step's
shapes: general functors against the lean row kernels (BN backward only), with / without reversed sweeps.
Buffers rotate over four sets (> the 126 MB L2 in total) so that one iteration does not find the previous one's data.
  python tools/exp/ew_bench.py [--iters 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mmhand_b200 import runtime  # noqa: E402
from mmhand_b200.kernels import GradSource  # noqa: E402
from mmhand_b200.layouts import geom_s1, geom_s2, geom_up  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--sets", type=int, default=4)
a = ap.parse_args()
ops = runtime.get_ops(torch.device("cuda", 0))
B = 16
CONFIGS = (("general", dict(MMH_EW_LEAN="0")),
           ("lean", dict(MMH_EW_LEAN="1", MMH_EW_REVERSE="0")),
           ("lean+rev", dict(MMH_EW_LEAN="1", MMH_EW_REVERSE="1")))


def shapes():
    # (name, xl (producer's raw output), consumer in_lay, lo, hi, reflect, relu, dropout)
    for Cc in (256, 512):
        gp, gc = geom_s1(B, 64, 64, 3, 'reflect', Cc, Cc), geom_s1(B, 64, 64, 3, 'reflect', Cc, Cc)
        yield "3x3 64^2 C=%d" % Cc, gp.out_lay, gc.in_lay, 1, 1, True, True, True
    gp, gc = geom_s1(B, 256, 256, 7, 'reflect', 64, 64), geom_s2(B, 256, 256, 64, 64)
    yield "7x7 256^2 C=64 -> s2", gp.out_lay, gc.in_lay, 1, 1, False, True, False
    gp, gc = geom_up(B, 128, 128, 64, 64), geom_s1(B, 256, 256, 7, 'reflect', 64, 64)
    yield "up 256^2 C=64 -> 7x7", gp.out_lay, gc.in_lay, 3, 3, True, True, False


def timed(fn, n):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000.0


for name, xl, sl, lo, hi, refl, relu, drop in shapes():
    Cc = xl.C
    S = a.sets
    mk = lambda rows, ld: [torch.randn(rows, ld, device="cuda").to(torch.bfloat16) for _ in range(S)]
    xs, srcs, dys, dsts = mk(xl.rows, xl.ld), mk(sl.rows, sl.ld), mk(xl.rows, xl.ld), mk(sl.rows, sl.ld)
    coef, save = torch.rand(2 * Cc, device="cuda") + 0.5, torch.rand(2 * Cc, device="cuda") + 0.5
    sums, k = torch.zeros(2 * Cc, device="cuda"), torch.zeros(2 * Cc, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    n_el = B * xl.H * xl.W * Cc
    halo = sl.rows * sl.ld / float(n_el)

    def reduce(i):
        j = i % S
        ops.bn_bwd_reduce_finalize(None, ([GradSource(srcs[j], sl, lo, hi, refl)], None), False, relu, drop, 0x1234,
                                   xs[j], xl, coef, save, sums, k, ticket, float(B * xl.H * xl.W), None, None)

    def apply(i):
        j = i % S
        ops.bn_bwd_apply(([GradSource(srcs[j], sl, lo, hi, refl)], None), False, relu, drop, 0x1234, xs[j], xl, coef,
                         save, k, dys[j], xl)

    def pair(i):
        reduce(i)
        apply(i)

    def norm(i):
        j = i % S
        ops.norm_act(xs[j], xl, coef, relu, drop, 0x1234, dsts[j], sl, lo, hi, refl)

    ref = None
    for cname, env in CONFIGS:
        os.environ.update(env)
        t_r, t_a, t_p, t_n = timed(reduce, a.iters), timed(apply, a.iters), timed(pair, a.iters), timed(norm, a.iters)
        # algorithmic bytes (bf16): reduce 4 / element, apply 6, norm 2 read + 2 written (+ halo)
        gb = lambda by, us: by * n_el / us / 1e3
        print("%-22s %-12s reduce %6.1f us %5.0f GB/s | apply %6.1f us %5.0f GB/s | pair %6.1f us | norm_act %6.1f us %5.0f GB/s"
              % (name, cname, t_r, gb(4, t_r), t_a, gb(6, t_a), t_p, t_n, gb(2 + 2 * halo, t_n)), flush=True)
        # results agree with the general path
        sums.zero_()
        pair(0)
        norm(0)
        torch.cuda.synchronize()
        got = (k.clone(), dys[0].float().clone(), dsts[0].float().clone())
        if ref is None:
            ref = got
        else:
            dk = (got[0] - ref[0]).abs().max().item() / max(ref[0].abs().max().item(), 1e-9)
            assert dk < 1e-4, (cname, "k", dk)
            assert (got[1] - ref[1]).abs().max().item() <= 1e-2 * ref[1].abs().max().item(), (cname, "dy")
            assert torch.equal(got[2], ref[2]), (cname, "norm_act")
    del xs, srcs, dys, dsts
