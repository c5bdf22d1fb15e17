"""Generate tests/golden/*.pt from the REAL reference modules (run in the build container only):

    python oracle/make_golden.py

Small configurations (ngf = ndf = 4, 32 x 32 frames) keep the fixtures to ~2 MB. The fixtures pin the oracle
(oracle/patn_ref.py) on boxes where /root/reference is absent.
"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synth_batch(B, S, gen):
    r = lambda *s: torch.rand(*s, generator=gen)
    return dict(H1=r(B, 3, S, S) * 2 - 1, P1=r(B, 21, S, S), D1=r(B, 3, S, S) * 2 - 1,
                H2=r(B, 3, S, S) * 2 - 1, P2=r(B, 21, S, S), D2=r(B, 3, S, S) * 2 - 1,
                H1_path=["a"] * B, H2_path=["b"] * B)


def main():
    os.makedirs(OUT, exist_ok=True)
    G, D, nu, Pool, L1P = ref_shims.load_reference_nets()
    gen = torch.Generator().manual_seed(49)
    torch.manual_seed(49)
    S, B, ngf = 32, 2, 4
    norm = nu.get_norm_layer('batch')
    # --- networks: forward passes
    g = G([3, 42, 6], 3, ngf, norm, True, 9)
    nu.init_weights(g, 'normal')
    d = D(24, ngf, norm, True, 3, [], 'reflect', False, 2)
    nu.init_weights(d, 'normal')
    x = [torch.rand(B, 3, S, S, generator=gen) * 2 - 1, torch.rand(B, 42, S, S, generator=gen),
         torch.rand(B, 6, S, S, generator=gen) * 2 - 1]
    xd = torch.rand(B, 24, S, S, generator=gen) * 2 - 1
    out = {"ngf": ngf, "x": x, "xd": xd}
    out["g_sd"] = {k: v.clone() for k, v in g.state_dict().items()}
    out["d_sd"] = {k: v.clone() for k, v in d.state_dict().items()}
    g.eval()
    d.eval()
    with torch.no_grad():
        out["g_eval"] = g(x)
        out["d_eval"] = d(xd)
    # train mode with dropout disabled through p=0 modules is not expressible; build no-dropout twins
    g2 = G([3, 42, 6], 3, ngf, norm, False, 9)
    nu.init_weights(g2, 'normal')
    d2 = D(24, ngf, norm, False, 3, [], 'reflect', False, 2)
    nu.init_weights(d2, 'normal')
    out["g2_sd"] = {k: v.clone() for k, v in g2.state_dict().items()}
    out["d2_sd"] = {k: v.clone() for k, v in d2.state_dict().items()}
    g2.train()
    d2.train()
    with torch.no_grad():
        out["g2_train"] = g2(x)
        out["d2_train"] = d2(xd)
    out["g2_sd_after"] = {k: v.clone() for k, v in g2.state_dict().items() if "running" in k}
    # --- losses
    crit = nu.GANLoss(use_lsgan=False, gpu='cpu')
    out["gan_real"] = crit(out["d2_train"], True)
    out["gan_fake"] = crit(out["d2_train"], False)
    torch.save(out, os.path.join(OUT, "nets_ngf4.pt"))

    # --- full training steps with the reference MMHandModel (dropout off: masks come from torch's RNG)
    MM = ref_shims.load_reference_model_class()
    torch.manual_seed(49)
    random.seed(49)
    opt = ref_shims.make_opt(batchSize=B, fineSize=S, ngf=ngf, ndf=ngf, no_dropout=True, no_dropout_D=True,
                             pool_size=3)
    m = MM(opt)
    step = {"opt": vars(opt).copy(), "sd_g": {k: v.clone() for k, v in m.netG.state_dict().items()},
            "sd_dpb": {k: v.clone() for k, v in m.netD_PB.state_dict().items()},
            "sd_dpp": {k: v.clone() for k, v in m.netD_PP.state_dict().items()},
            "sd_vgg": {k: v.clone() for k, v in m.criterionL1.vgg_submodel.state_dict().items()},
            "batches": [], "errors": [], "fake": []}
    for it in range(4):
        b = synth_batch(B, S, gen)
        step["batches"].append({k: v for k, v in b.items() if not k.endswith("path")})
        m.set_input(b)
        m.optimize_parameters()
        step["errors"].append({k: float(v) for k, v in m.get_current_errors().items()})
        step["fake"].append(m.fake_p2.detach().clone())
    step["final_g_sum"] = {k: float(v.double().abs().sum()) for k, v in m.netG.state_dict().items()
                           if v.is_floating_point()}
    step["final_dpb_sum"] = {k: float(v.double().abs().sum()) for k, v in m.netD_PB.state_dict().items()
                             if v.is_floating_point()}
    torch.save(step, os.path.join(OUT, "step_ngf4.pt"))
    # norm='instance' variants of the reference's own modules: checkpoint ABI only (tests/test_state_dict_compat.py)
    RG, RD, nu, _, _ = ref_shims.load_reference_nets()
    inorm = nu.get_norm_layer('instance')
    torch.save({"g_in_sd": RG([3, 42, 6], 3, 4, inorm, True, 9).state_dict(),
                "d_in_sd": RD(24, 4, inorm, True, 3, [], 'reflect', False, 2).state_dict()},
               os.path.join(OUT, "nets_instance_ngf4.pt"))
    print("errors:", step["errors"])
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
