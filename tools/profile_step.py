"""One G+D training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 16]
"""
import argparse
import contextlib
import io
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import synth_batch  # noqa: E402
from models.MMHandModel import MMHandModel  # noqa: E402
from mmhand_b200.options import make_opt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
torch.cuda.set_device(0)
torch.manual_seed(49)
random.seed(49)
opt = make_opt(batchSize=a.batch, fineSize=a.size, local_rank=0, gpu=0, seed=49)
with contextlib.redirect_stdout(io.StringIO()):
    m = MMHandModel(opt)
b = {k: v.cuda() for k, v in synth_batch(a.batch, a.size, 1).items()}
for _ in range(a.warmup):
    m.set_input(b)
    m.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    m.set_input(b)
    m.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
