"""Diagnostic (GPU): run the compact-input test's step sequence with guard gaps behind every arena buffer
(MMH_ARENA_GUARD) and report out-of-bounds writes / the first step whose losses are not finite."""
import os
import random
import sys

os.environ.setdefault("MMH_ARENA_GUARD", "4096")
os.environ.setdefault("MMH_VGG19_RANDOM", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import math  # noqa: E402

import torch  # noqa: E402

from mmhand_b200 import runtime  # noqa: E402
from mmhand_b200.options import make_opt  # noqa: E402
from models.MMHandModel import MMHandModel  # noqa: E402

B, S = int(os.environ.get("DIAG_B", "2")), 256
n_models, n_steps = int(os.environ.get("DIAG_MODELS", "4")), int(os.environ.get("DIAG_STEPS", "4"))
for mi in range(n_models):
    torch.manual_seed(4)
    random.seed(4)
    m = MMHandModel(make_opt(batchSize=B, fineSize=S, pool_size=0, local_rank=0, gpu=0, seed=3))
    m.master = False
    ops = runtime.get_ops(torch.device("cuda", 0))
    g = torch.Generator().manual_seed(90)
    r = lambda *s: torch.rand(*s, generator=g)
    for it in range(n_steps):
        b = dict(H1=r(B, 3, S, S) * 2 - 1, P1=(r(B, 21, S, S) > 0.98).float() * r(B, 21, S, S), D1=r(B, 3, S, S) * 2 - 1,
                 H2=r(B, 3, S, S) * 2 - 1, P2=(r(B, 21, S, S) > 0.98).float() * r(B, 21, S, S), D2=r(B, 3, S, S) * 2 - 1)
        m.set_input(b)
        m.optimize_parameters()
        errs = {k: float(v) for k, v in m.get_current_errors().items()}
        torch.cuda.synchronize()
        bad = ops.check_guards()
        ok = all(math.isfinite(v) for v in errs.values())
        print("model %d step %d finite=%s L1=%.4f D_PP=%.4f guards_dirty=%d" % (mi, it, ok, errs["pair_L1loss"],
                                                                              errs["D_PP"], len(bad)), flush=True)
        for line in bad[:12]:
            print("    ", line, flush=True)
        if bad or not ok:
            # which parameters / statistics are not finite
            for name, net in (("G", m.netG), ("D_PB", m.netD_PB), ("D_PP", m.netD_PP)):
                nf = [k for k, v in net.state_dict().items() if v.is_floating_point() and not torch.isfinite(v).all()]
                if nf:
                    print("     non-finite in %s: %d tensors, first %s" % (name, len(nf), nf[:4]), flush=True)
            break
    del m
    torch.cuda.empty_cache()
