"""Per-step device time of the training step (CUDA events per step), device-resident and host-fed inputs, with and
without the nvidia-smi clock sampler thread -- diagnostic for bench.py's timed regions."""
import contextlib
import io
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import ClockSampler, synth_batch  # noqa: E402
from models.MMHandModel import MMHandModel  # noqa: E402
from mmhand_b200.options import make_opt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.cuda.set_device(0)
torch.manual_seed(49)
random.seed(49)
opt = make_opt(batchSize=B, fineSize=256, local_rank=0, gpu=0, seed=49)
with contextlib.redirect_stdout(io.StringIO()):
    m = MMHandModel(opt)
host = [synth_batch(B, 256, 1000 + i, pin=True) for i in range(2)]
dev = [{k: v.cuda() for k, v in h.items()} for h in host]


def run(n, feed, tag, read_back=False):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    hs = []
    torch.cuda.synchronize()
    t0 = time.time()
    evs[0].record()
    for i in range(n):
        h0 = time.time()
        m.set_input(feed[i % 2])
        m.optimize_parameters()
        if read_back:
            torch.stack([v.reshape(()) for v in m.get_current_errors().values()]).cpu()
        evs[i + 1].record()
        hs.append((time.time() - h0) * 1e3)
    torch.cuda.synchronize()
    wall = (time.time() - t0) * 1e3
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    print("%-28s wall %.1f ms/step | device per step: %s" % (tag, wall / n, " ".join("%.1f" % x for x in ms)))
    print("%-28s host-side issue ms per step: %s" % ("", " ".join("%.1f" % x for x in hs)))


run(3, dev, "warmup")
run(12, dev, "dev inputs")
run(12, host, "host inputs + readback", True)
run(12, dev, "dev inputs again")
s = ClockSampler(0)
s.start()
run(12, dev, "dev inputs + clock sampler")
run(12, host, "host inputs + sampler", True)
s.stop_flag = True
s.join(timeout=3)
print(s.summary())
