"""Diagnostic (CPU, oracle only): which bf16 storage point of the generator breaks the north-star's 2e-2 max-abs?

The fp32 oracle is re-run with bf16 rounding at a subset of the points where the CUDA path stores bf16
(conv input activations "act", weights "w", raw conv outputs "raw") and compared with itself in plain fp32.
Output: profiles/r02_quant_ablation.txt (VERDICT r1 "Next round" item 1a)."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import patn_ref as O  # noqa: E402
from models.Generator import Generator  # noqa: E402
from models.network_utils import get_norm_layer, init_weights  # noqa: E402


def main():
    B, S = 2, 256
    torch.manual_seed(49)
    g = Generator([3, 42, 6], 3, 64, get_norm_layer('batch'), True, 9)
    init_weights(g, 'normal')
    sd = {k: v.detach().clone() for k, v in g.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    r = lambda *s: torch.rand(*s, generator=gen)
    x = [r(B, 3, S, S) * 2 - 1, (r(B, 42, S, S) > 0.98).float() * r(B, 42, S, S), r(B, 6, S, S) * 2 - 1]
    rows = []
    for train in (True, False):
        sdx = {k: v.clone() for k, v in sd.items()}
        if not train:
            O.BN_MOM = 1.0
            with torch.no_grad():
                O.generator_forward(sdx, x, train=True, use_dropout=True)
            O.BN_MOM = 0.1
        run = lambda: O.generator_forward({k: v.clone() for k, v in sdx.items()}, x, train=train, use_dropout=True,
                                          drop=O.DropCtx("hash", 0, 0, 0))
        with torch.no_grad():
            ref = run()
            O.QUANT = lambda t: t.bfloat16().float()
            for n in (1, 2, 3):
                for pts in itertools.combinations(("act", "w", "raw"), n):
                    O.QUANT_POINTS = set(pts)
                    y = run()
                    e = (y - ref).abs()
                    rows.append(("train" if train else "eval", "+".join(pts), e.max().item(), e.mean().item()))
                    print(rows[-1], flush=True)
            O.QUANT, O.QUANT_POINTS = None, None
    with open(os.path.join(ROOT, "profiles", "r02_quant_ablation.txt"), "w") as f:
        f.write("generator output (tanh, [-1,1]) ngf=64 256x256 B=2: fp32 oracle vs the same oracle with bf16 rounding at\n"
                "the listed storage points (fp32 accumulation / statistics / trunk everywhere)\n")
        f.write("%-6s %-12s %10s %10s\n" % ("mode", "bf16 points", "max-abs", "mean-abs"))
        for row in rows:
            f.write("%-6s %-12s %10.4f %10.5f\n" % row)


if __name__ == "__main__":
    main()
