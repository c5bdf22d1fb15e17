"""Per-layer device time of every conv / wgrad launch of one eager training step (CUDA events per launch)."""
import argparse
import contextlib
import io
import os
import random
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import synth_batch  # noqa: E402
from mmhand_b200 import runtime  # noqa: E402
from models.MMHandModel import MMHandModel  # noqa: E402
from mmhand_b200.options import make_opt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
torch.cuda.set_device(0)
torch.manual_seed(49)
random.seed(49)
opt = make_opt(batchSize=a.batch, fineSize=256, local_rank=0, gpu=0, seed=49)
with contextlib.redirect_stdout(io.StringIO()):
    m = MMHandModel(opt)
m.use_tape = False
b = {k: v.cuda() for k, v in synth_batch(a.batch, 256, 1).items()}
for _ in range(2):
    m.set_input(b)
    m.optimize_parameters()
ops = runtime.get_ops(torch.device("cuda", 0))
recs = []


def hook(kind, tag, plan, launch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    d = plan.desc
    if kind == "conv":
        fl = 2.0 * d.M * d.N * min(d.C, 10 ** 9) * d.T * (d.Hv * d.Wv) / float(d.Hg * d.Wg)
        shape = "M=%d N=%d C=%d T=%d" % (d.M, d.N, d.C, d.T)
    else:
        fl = 2.0 * d.M * d.N * d.C * d.T
        shape = "M=%d N=%d C=%d T=%d" % (d.M, d.N, d.C, d.T)
    recs.append((tag, kind, shape, fl, e0, e1))


ops.conv_hook = hook
m.set_input(b)
m.optimize_parameters()
torch.cuda.synchronize()
ops.conv_hook = None
agg = OrderedDict()
for tag, kind, shape, fl, e0, e1 in recs:
    name = tag[0] if tag else "?"
    # collapse block indices: b3.s1.c1 -> b*.s1.c1 ; r2.c1 -> r*.c1
    import re
    key = (re.sub(r"^b[1-8]\.", "b1-8.", re.sub(r"^r\d\.", "r*.", name)), tag[1] if tag else kind, shape)
    v = agg.setdefault(key, [0, 0.0, 0.0])
    v[0] += 1
    v[1] += e0.elapsed_time(e1)
    v[2] += fl
tot = sum(v[1] for v in agg.values())
print("%-16s %-6s %-34s %5s %9s %8s %8s" % ("layer", "pass", "shape", "n", "ms", "share", "TFLOP/s"))
for (name, kind, shape), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-16s %-6s %-34s %5d %9.3f %7.1f%% %8.1f" % (name, kind, shape, n, ms, 100 * ms / tot, fl / ms / 1e9))
print("TOTAL conv+wgrad ms %.3f  (%.1f TFLOP padded)" % (tot, sum(v[2] for v in agg.values()) / 1e12))
