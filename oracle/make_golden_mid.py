"""Generate tests/golden/mid_outputs.pt from the REAL reference modules (run in the build container only):

    python oracle/make_golden_mid.py

A mid-size configuration -- ngf = ndf = 16, 64 x 64, 9 PAT blocks / 3 residual blocks, batch 2 -- pinned by its OUTPUTS
only: the weights come from oracle/golden_weights.fill (a function of seed, key and shape), so the test rebuilds them
without the reference. Eval mode (running statistics) and train mode (batch statistics; dropout layers absent so that no
RNG is involved) for the generator, the discriminator and the two-stream PATNetwork variant.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from oracle.golden_weights import fill  # noqa: E402


def inputs(seed=33, S=64, B=2):
    gen = torch.Generator().manual_seed(seed)
    x = [torch.rand(B, 3, S, S, generator=gen) * 2 - 1, torch.rand(B, 42, S, S, generator=gen),
         torch.rand(B, 6, S, S, generator=gen) * 2 - 1]
    xd = torch.rand(B, 24, S, S, generator=gen) * 2 - 1
    return x, xd


def main():
    G, D, nu, Pool, L1P = ref_shims.load_reference_nets()
    norm = nu.get_norm_layer('batch')
    nf = 16
    x, xd = inputs()
    out = {"input_seed": 33}          # the inputs are regenerated from this seed (same draws) by the test
    g = G([3, 42, 6], 3, nf, norm, False, 9)
    g.load_state_dict(fill(g.state_dict(), 1))
    d = D(24, nf, norm, False, 3, [], 'reflect', False, 2)
    d.load_state_dict(fill(d.state_dict(), 2))
    with torch.no_grad():
        g.eval(); d.eval()
        out["g_eval"], out["d_eval"] = g(x), d(xd)
        g.train(); d.train()
        out["g_train"], out["d_train"] = g(x), d(xd)
    torch.save(out, os.path.join(ROOT, "tests", "golden", "mid_outputs.pt"))
    print({k: (tuple(v.shape), float(v.abs().mean())) for k, v in out.items() if torch.is_tensor(v)})


if __name__ == "__main__":
    main()
