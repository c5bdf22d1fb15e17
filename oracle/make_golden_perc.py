"""Generate tests/golden/perc_layers.pt from the REAL reference loss class (run in the build container only):

    python oracle/make_golden_perc.py

losses/L1_plus_perceptualLoss.py:22-27 cuts ``vgg19.features`` after index ``perceptual_layers``; the shipped value is 3.
The fixture pins oracle.patn_ref.l1_plus_perceptual(perceptual_layers = 0, 1, 2, 3) (L1 and MSE flavours) on 2 x 3 x 24 x
24 images with the random-init VGG19 of the shims (no network here).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402


def main():
    ref_shims.load_reference_model_class()          # installs the vgg19 / DataParallel / .cuda() shims
    _, _, _, _, L1P = ref_shims.load_reference_nets()
    g = torch.Generator().manual_seed(21)
    x = torch.rand(2, 3, 24, 24, generator=g) * 2 - 1
    t = torch.rand(2, 3, 24, 24, generator=g) * 2 - 1
    out = {"x": x, "t": t}
    for p in (0, 1, 2, 3):
        for is_l1 in (1, 0):
            torch.manual_seed(100 + p)
            crit = L1P(10.0, 10.0, p, ['cpu'], is_l1)
            xi = x.clone().requires_grad_(True)
            loss, l1, lp = crit(xi, t)
            loss.backward()
            out["p%d_l1%d" % (p, is_l1)] = dict(sd={k: v.clone() for k, v in crit.vgg_submodel.state_dict().items()},
                                                loss=loss.detach(), l1=l1.detach(), lp=lp.detach(), grad=xi.grad.clone())
    torch.save(out, os.path.join(ROOT, "tests", "golden", "perc_layers.pt"))
    print({k: (float(v["l1"]), float(v["lp"])) for k, v in out.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
