"""Diagnostic: per-stage error of the CUDA generator against the fp32 oracle and against the oracle under bf16
autocast (what stock PyTorch bf16 gives) -- shows how much of the deviation is inherent to bf16 storage."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from models.Generator import Generator  # noqa: E402
from models.network_utils import get_norm_layer, init_weights  # noqa: E402
from oracle import patn_ref as O  # noqa: E402

torch.manual_seed(49)
dev = "cuda"
g = Generator([3, 42, 6], 3, 64, get_norm_layer('batch'), True, 9).to(dev)
init_weights(g, 'normal')
gen = torch.Generator().manual_seed(1)
r = lambda *s: torch.rand(*s, generator=gen).to(dev)
B, S = 2, 256
x = [r(B, 3, S, S) * 2 - 1, (r(B, 42, S, S) > 0.98).float() * r(B, 42, S, S), r(B, 6, S, S) * 2 - 1]
sd = {k: v.detach().clone() for k, v in g.state_dict().items()}
g.train()
g._step = 0
with torch.no_grad():
    y = g(x)
    t32, t16 = {}, {}
    y32 = O.generator_forward({k: v.clone() for k, v in sd.items()}, x, train=True, use_dropout=True,
                              drop=O.DropCtx("hash", 0, 0, 0), taps=t32)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = O.generator_forward({k: v.clone() for k, v in sd.items()}, x, train=True, use_dropout=True,
                                  drop=O.DropCtx("hash", 0, 0, 0), taps=t16).float()
eng = g.engine(B, S, S)
print("output: ours-vs-fp32 max %.4f mean %.5f | torch-bf16-autocast-vs-fp32 max %.4f mean %.5f" % (
    (y - y32).abs().max(), (y - y32).abs().mean(), (y16 - y32).abs().max(), (y16 - y32).abs().mean()))
for i in range(9):
    a, b = t32["att%d" % i], t16["att%d" % i].float()
    print("block %d out: scale %.3f  torch-bf16 err max %.4f mean %.5f" % (i, a.abs().mean(), (a - b).abs().max(),
                                                                         (a - b).abs().mean()))
# our trunk after the last block (fp32 NHWC plain buffer)
cur = eng.trunk[9 % 2].view(B, S // 4, S // 4, -1).permute(0, 3, 1, 2)
a = t32["att8"]
print("ours block 8 out err max %.4f mean %.5f" % ((a - cur).abs().max(), (a - cur).abs().mean()))
