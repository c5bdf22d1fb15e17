"""Generate tests/golden/disc_variants_ngf4.pt from the REAL reference Discriminator (run in the build container only):

    python oracle/make_golden_dvariants.py

The reference class accepts n_downsampling in {0, 1, 2, 3} (models/Discriminator.py:86-133; 3 adds a 4 ndf -> 4 ndf
stride-2 stage). The fixture pins oracle.patn_ref.discriminator_forward(n_downsampling=...) for the values the shipped
configuration does not use (eval-mode forward, no dropout randomness: ngf = 4, 32 x 32, two residual blocks).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402


def main():
    G, D, nu, Pool, L1P = ref_shims.load_reference_nets()
    gen = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    out = {}
    norm = nu.get_norm_layer('batch')
    for nd in (1, 3):
        d = D(6, 4, norm, True, 2, [], 'reflect', False, nd)
        nu.init_weights(d, 'normal')
        x = torch.rand(2, 6, 32, 32, generator=gen) * 2 - 1
        d.train()
        with torch.no_grad():
            d(x)                                   # moves the running statistics off their initial values
        d.eval()
        with torch.no_grad():
            y = d(x)
        out["nd%d" % nd] = dict(sd={k: v.clone() for k, v in d.state_dict().items()}, x=x, y=y)
    torch.save(out, os.path.join(ROOT, "tests", "golden", "disc_variants_ngf4.pt"))
    print({k: tuple(v["y"].shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
