"""Per-entry-point device time of one training step measured with CUDA events around every launch of an eager
(un-taped) step: warm L2 and real predecessor kernels, unlike ncu's serialised cold-cache replays. Diagnostic:
shows where L2 residency between producer and consumer kernels already helps and where it does not."""
import collections
import contextlib
import io
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import synth_batch  # noqa: E402
from mmhand_b200 import runtime  # noqa: E402
from mmhand_b200.options import make_opt  # noqa: E402
from models.MMHandModel import MMHandModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.cuda.set_device(0)
torch.manual_seed(49)
random.seed(49)
opt = make_opt(batchSize=B, fineSize=256, local_rank=0, gpu=0, seed=49)
with contextlib.redirect_stdout(io.StringIO()):
    m = MMHandModel(opt)
dev = [{k: v.cuda() for k, v in synth_batch(B, 256, 1000 + i).items()} for i in range(2)]
for i in range(4):
    m.set_input(dev[i % 2])
    m.optimize_parameters()
ops = runtime.get_ops(torch.device("cuda", 0))
recs = []
orig = ops._run


def timed_run(fn, args, patch=None, keep=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = ops.in_side if ops.in_side is not None else torch.cuda.current_stream()
    e0.record(st)
    orig(fn, args, patch, keep)
    e1.record(st)
    recs.append((fn.__name__, e0, e1))


m.use_tape = False
ops._run = timed_run
m.set_input(dev[0])
m.optimize_parameters()
torch.cuda.synchronize()
ops._run = orig
tot = collections.defaultdict(lambda: [0, 0.0])
for name, e0, e1 in recs:
    tot[name][0] += 1
    tot[name][1] += e0.elapsed_time(e1)
total = sum(v[1] for v in tot.values())
print("%-28s %8s %10s %7s" % ("entry point", "launches", "ms", "share"))
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %8d %10.3f %6.1f%%" % (name, n, ms, 100.0 * ms / total))
print("%-28s %8d %10.3f" % ("TOTAL (event-bracketed)", len(recs), total))
