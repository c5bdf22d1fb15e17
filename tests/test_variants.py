"""the two-stream pose-transfer generator of the benchmark harness (networks/model_variants.py) and the
evaluator's SSIM -- oracle pinned by golden vectors from the reference's own classes (oracle/make_golden_variants.py),
host logic of the engine (two layer chains, one attention map, no swap) against the oracle on the host emulation."""
import os

import pytest
import torch

import hostemu
from mmhand_b200 import runtime
from oracle import patn_ref as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "patn2_ngf4.pt")


@pytest.fixture
def emu():
    runtime._TEST_OPS = hostemu.ops(f32=True)
    yield
    runtime._TEST_OPS = None


def test_oracle_reproduces_the_reference_two_stream_network_and_ssim():
    g = torch.load(GOLD)
    with torch.no_grad():
        y = O.generator2_forward(g["sd"], g["x"], train=False)
        assert torch.allclose(y, g["eval"], atol=1e-6), (y - g["eval"]).abs().max()
        sd2 = {k: v.clone() for k, v in g["sd2"].items()}
        y2 = O.generator2_forward(sd2, g["x"], train=True, use_dropout=False)
        assert torch.allclose(y2, g["train"], atol=1e-6)
        for k, v in g["sd2_after"].items():
            assert torch.allclose(sd2[k], v, atol=1e-6), k
    assert abs(float(O.ssim(g["ssim_a"], g["ssim_b"])) - float(g["ssim_mean"])) < 1e-7
    assert torch.allclose(O.ssim(g["ssim_a"], g["ssim_b"], size_average=False), g["ssim_per_image"], atol=1e-7)


def test_state_dict_keys_are_the_reference_classes():
    from models.network_utils import get_norm_layer
    from networks.model_variants import PATNetwork
    g = torch.load(GOLD)
    net = PATNetwork([3, 3], 3, g["ngf"], get_norm_layer('batch'), True, 9)
    mine = net.state_dict()
    assert list(mine.keys()) == list(g["sd"].keys())
    assert all(mine[k].shape == g["sd"][k].shape and mine[k].dtype == g["sd"][k].dtype for k in mine)
    net.load_state_dict(g["sd"])
    with pytest.raises(AssertionError):
        PATNetwork([3, 3, 3], 3)


def test_two_stream_engine_forward_backward(emu):
    from models.network_utils import get_norm_layer, init_weights
    from networks.model_variants import PATNetwork
    torch.manual_seed(7)
    net = PATNetwork([3, 21], 3, 16, get_norm_layer('batch'), True, 9)
    init_weights(net, 'normal')
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(8)
    x = [torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1, torch.rand(2, 21, 32, 32, generator=gen)]
    net.train()
    y = net(x)
    gy = torch.randn(y.shape, generator=gen)
    y.backward(gy)
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    want = O.generator2_forward(sdo, x, train=True, use_dropout=True, drop=O.DropCtx("hash", 0, 0, 0))
    want.backward(gy)
    assert torch.allclose(y.detach(), want.detach(), atol=5e-5), (y.detach() - want.detach()).abs().max()
    for k, p in net.named_parameters():
        r = sdo[k].grad
        assert (p.grad - r).abs().max() <= 2e-4 * r.abs().max() + 1e-7, k
    # eval mode on the golden weights: the reference class's own output
    g = torch.load(GOLD)
    net2 = PATNetwork([3, 3], 3, 16, get_norm_layer('batch'), True, 9)
    init_weights(net2, 'normal')
    sd2 = {k: v.detach().clone() for k, v in net2.state_dict().items()}
    net2.eval()
    with torch.no_grad():
        x2 = [t[:, :3] for t in (x[0], x[1])]
        assert torch.allclose(net2(x2), O.generator2_forward(sd2, x2, train=False), atol=2e-5)


def test_ssim_kernel(emu):
    from mmhand_b200.metrics import ssim
    g = torch.load(GOLD)
    a, b = g["ssim_a"], g["ssim_b"]
    assert abs(float(ssim(a, b)) - float(g["ssim_mean"])) < 2e-6
    assert torch.allclose(ssim(a, b, size_average=False).cpu(), g["ssim_per_image"], atol=2e-6)
    assert abs(float(ssim(a, a)) - float(g["ssim_same"])) < 2e-6
    odd = torch.rand(1, 1, 7, 5)
    assert abs(float(ssim(odd, odd * 0.5)) - float(O.ssim(odd, odd * 0.5))) < 2e-6          # frame smaller than the window
