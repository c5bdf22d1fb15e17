"""The oracle (oracle/patn_ref.py) against golden vectors produced by the real reference modules
(oracle/make_golden.py): networks, losses and four full optimisation steps."""
import os
import random

import pytest
import torch

from oracle import patn_ref as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def nets():
    return torch.load(os.path.join(GOLD, "nets_ngf4.pt"))


def test_generator_eval(nets):
    y = O.generator_forward(nets["g_sd"], nets["x"], train=False, use_dropout=True)
    assert torch.allclose(y, nets["g_eval"], atol=1e-6)
    assert y.shape == (2, 3, 32, 32)


def test_generator_train_batchnorm(nets):
    sd = {k: v.clone() for k, v in nets["g2_sd"].items()}
    y = O.generator_forward(sd, nets["x"], train=True, use_dropout=False)
    assert torch.allclose(y, nets["g2_train"], atol=1e-5)
    for k, v in nets["g2_sd_after"].items():
        assert torch.allclose(sd[k], v, atol=1e-6), k


def test_discriminator(nets):
    y = O.discriminator_forward(nets["d_sd"], nets["xd"], train=False, use_dropout=True)
    assert torch.allclose(y, nets["d_eval"], atol=1e-5)
    sd = {k: v.clone() for k, v in nets["d2_sd"].items()}
    y2 = O.discriminator_forward(sd, nets["xd"], train=True, use_dropout=False)
    assert torch.allclose(y2, nets["d2_train"], atol=1e-5)
    assert y2.shape == (2, 16, 8, 8)     # no 1-channel head: logits are the last block's features (Q4)


def test_gan_loss(nets):
    assert abs(O.gan_loss(nets["d2_train"], True).item() - nets["gan_real"].item()) < 1e-6
    assert abs(O.gan_loss(nets["d2_train"], False).item() - nets["gan_fake"].item()) < 1e-6
    # known answer: BCE-with-logits of 0 against any label is ln 2
    assert abs(O.gan_loss(torch.zeros(2, 3, 4, 4), True).item() - 0.6931471805599453) < 1e-7


def test_swap_quirk(nets):
    """Q1: perturbing only the pose input must reach block 1 through conv_block_stream3, not stream2."""
    sd = nets["g_sd"]
    x = [t.clone() for t in nets["x"]]
    t0, t1 = {}, {}
    O.generator_forward(sd, x, train=False, taps=t0)
    x[1] = x[1] + 0.5
    O.generator_forward(sd, x, train=False, taps=t1)
    # the image-stream stem output is untouched, the block outputs are not
    assert torch.equal(t0["down"][0], t1["down"][0])
    assert not torch.equal(t0["att0"], t1["att0"])


def test_four_training_steps():
    g = torch.load(os.path.join(GOLD, "step_ngf4.pt"))
    o = g["opt"]
    random.seed(49)
    tr = O.OracleTrainer(g["sd_g"], g["sd_dpb"], g["sd_dpp"], g["sd_vgg"], o["lambda_A"], o["lambda_B"],
                         o["lambda_GAN"], o["lr"], o["beta1"], o["pool_size"], use_dropout_g=False,
                         use_dropout_d=False)
    for b, want, fake in zip(g["batches"], g["errors"], g["fake"]):
        got = tr.step(b["H1"], b["P1"], b["D1"], b["H2"], b["P2"], b["D2"])
        assert torch.allclose(tr.fake, fake, atol=2e-5)
        for k in want:
            assert abs(got[k] - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (k, got[k], want[k])
    for k, s in g["final_g_sum"].items():
        assert abs(float(tr.g[k].double().abs().sum()) - s) <= 1e-4 * max(1.0, s), k
    for k, s in g["final_dpb_sum"].items():
        assert abs(float(tr.dpb[k].double().abs().sum()) - s) <= 1e-4 * max(1.0, s), k


def test_discriminator_variants_match_reference_class():
    """oracle.discriminator_forward(n_downsampling = 1, 3) reproduces the reference Discriminator (eval mode) recorded by
    oracle/make_golden_dvariants.py."""
    import os
    import torch
    from oracle import patn_ref as O
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "disc_variants_ngf4.pt"))
    for nd in (1, 3):
        c = g["nd%d" % nd]
        y = O.discriminator_forward(c["sd"], c["x"], train=False, use_dropout=True, n_blocks=2, n_downsampling=nd)
        assert y.shape == c["y"].shape
        assert torch.allclose(y, c["y"], atol=1e-6), (nd, (y - c["y"]).abs().max())


def test_perceptual_layer_variants_match_reference_class():
    """oracle.l1_plus_perceptual(perceptual_layers = 0..3, L1 and MSE) reproduces the reference L1_plus_perceptualLoss
    (values and input gradient) recorded by oracle/make_golden_perc.py."""
    import os
    import torch
    from oracle import patn_ref as O
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "perc_layers.pt"))
    for p in (0, 1, 2, 3):
        for is_l1 in (1, 0):
            c = g["p%d_l1%d" % (p, is_l1)]
            x = g["x"].clone().requires_grad_(True)
            loss, l1, lp = O.l1_plus_perceptual(c["sd"], x, g["t"], 10.0, 10.0, is_l1, perceptual_layers=p)
            loss.backward()
            assert abs(float(l1) - float(c["l1"])) < 1e-5 and abs(float(lp) - float(c["lp"])) < 1e-5 * max(1, float(c["lp"]))
            assert torch.allclose(x.grad, c["grad"], atol=1e-7, rtol=1e-4), (p, is_l1)


def test_mid_size_outputs_match_reference_modules():
    """ngf = ndf = 16, 64 x 64, 9 PAT blocks / 3 residual blocks: the oracle reproduces the reference Generator /
    Discriminator outputs (eval and train mode) stored by oracle/make_golden_mid.py. Weights come from
    oracle/golden_weights.fill over the state_dict keys of THIS repository's drop-in modules (the reference's key set --
    tests/test_state_dict_compat.py), inputs from the recorded seed."""
    import os
    import torch
    from models.Discriminator import Discriminator
    from models.Generator import Generator
    from models.network_utils import get_norm_layer
    from oracle import patn_ref as O
    from oracle.golden_weights import fill
    from oracle.make_golden_mid import inputs
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mid_outputs.pt"))
    x, xd = inputs(g["input_seed"])
    norm = get_norm_layer('batch')
    sd_g = fill(Generator([3, 42, 6], 3, 16, norm, False, 9).state_dict(), 1)
    sd_d = fill(Discriminator(24, 16, norm, False, 3, [], 'reflect', False, 2).state_dict(), 2)
    with torch.no_grad():
        for mode, train in (("eval", False), ("train", True)):
            yg = O.generator_forward({k: v.clone() for k, v in sd_g.items()}, x, train=train, use_dropout=False)
            yd = O.discriminator_forward({k: v.clone() for k, v in sd_d.items()}, xd, train, False)
            assert torch.allclose(yg, g["g_" + mode], atol=2e-5), (mode, (yg - g["g_" + mode]).abs().max())
            assert torch.allclose(yd, g["d_" + mode], atol=2e-4, rtol=1e-4), (mode, (yd - g["d_" + mode]).abs().max())
