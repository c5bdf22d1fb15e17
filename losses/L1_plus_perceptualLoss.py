"""Drop-in for the reference's ``losses/L1_plus_perceptualLoss.py`` (:11-75):
``lambda_L1 * L1(inputs, targets) + lambda_perceptual * L1|MSE(VGG19[:perceptual_layers+1](n(inputs)), ...(n(targets)))``
with n(x) = ((x + 1) / 2 - mean) / std, returning ``(loss, loss_l1, loss_perceptual)``.

The VGG slice runs on the tcgen05 conv kernels (bias + ReLU in the epilogue), the two reductions and their
gradients on the fused loss kernels. ``perceptual_layers`` 0...3 are built (3 = conv1_1, ReLU, conv1_2, ReLU is the value
every shipped script uses; deeper slices would need the max-pool stages). Pretrained weights, in this order: ``MMH_VGG19_WEIGHTS`` (a torchvision ``vgg19``
state_dict file), torchvision's hub cache, ``torchvision.models.vgg19(weights=IMAGENET1K_V1)`` (a download, what the
reference does, :22). If none of them works the constructor RAISES -- a training run must not silently optimise against
random features -- unless ``MMH_VGG19_RANDOM=1`` opts into the seeded random initialisation (tests, benchmarks and the
smoke run: there is no network on the GPU boxes, and parity tests give both sides the same tensors).
"""
import os
import warnings

import torch
import torch.nn as nn

from mmhand_b200 import runtime
from mmhand_b200.engine import VggEngine


def _vgg_slice(perceptual_layers):
    """vgg19.features[0 : perceptual_layers + 1] as the reference builds it (:22-27)."""
    if not (0 <= perceptual_layers <= 3):
        raise NotImplementedError("perceptual_layers=%r: the B200 path builds VGG19.features up to index 3 (conv1_1, ReLU, "
                                  "conv1_2, ReLU; 3 is the value every shipped script uses) -- deeper slices need the "
                                  "max-pool stages" % (perceptual_layers,))
    layers = [("0", nn.Conv2d(3, 64, 3, padding=1)), ("1", nn.ReLU(inplace=True)), ("2", nn.Conv2d(64, 64, 3, padding=1)),
              ("3", nn.ReLU(inplace=True))][:perceptual_layers + 1]
    seq = nn.Sequential()
    for name, m in layers:
        seq.add_module(name, m)
    path = os.environ.get("MMH_VGG19_WEIGHTS", "")
    sd = None
    if path and os.path.exists(path):
        sd = torch.load(path, map_location="cpu")
    else:
        cache = os.path.expanduser("~/.cache/torch/hub/checkpoints/vgg19-dcbb9e9d.pth")
        if os.path.exists(cache):
            sd = torch.load(cache, map_location="cpu")
    random_ok = os.environ.get("MMH_VGG19_RANDOM", "0") == "1"
    if sd is None and not random_ok:
        try:                                    # the reference's own route (downloads into the hub cache)
            import torchvision
            sd = torchvision.models.vgg19(weights=torchvision.models.VGG19_Weights.IMAGENET1K_V1).state_dict()
        except Exception as e:                  # no network / no torchvision
            raise RuntimeError(
                "pretrained VGG19 weights are not available (%s: %s). Point MMH_VGG19_WEIGHTS at a torchvision vgg19 "
                "state_dict (vgg19-dcbb9e9d.pth), or set MMH_VGG19_RANDOM=1 to run the perceptual loss on seeded "
                "random features (tests / benchmarks only)." % (type(e).__name__, e)) from e
    if sd is not None:
        want = ["features.0.weight", "features.0.bias"] + (["features.2.weight", "features.2.bias"]
                                                           if perceptual_layers >= 2 else [])
        seq.load_state_dict({k[len("features."):]: v for k, v in sd.items() if k in want})
    else:
        warnings.warn("MMH_VGG19_RANDOM=1: the perceptual loss uses random-init VGG19 features")
    return seq


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, inputs, targets):
        x = inputs.contiguous().float()
        t = targets.contiguous().float()
        ops = runtime.get_ops(x.device)
        B, _, H, W = x.shape
        acc = torch.zeros(2, dtype=torch.float32, device=x.device)
        need = inputs.requires_grad
        g = torch.zeros_like(x) if need else None
        n = x.numel()
        ops.l1(x, t, mod.lambda_L1 / n, mod.lambda_L1 / n, acc[0:1], g)
        mod.vgg_engine(B, H, W).loss_and_backward(x, t, mod.lambda_perceptual, mod.percep_is_l1 != 1, acc[1:2], g)
        ctx.g = g
        ctx.mark_non_differentiable = None
        return acc[0] + acc[1], acc[0].clone(), acc[1].clone()

    @staticmethod
    def backward(ctx, g_loss, g_l1, g_p):
        if ctx.g is None:
            return None, None, None
        return None, ctx.g * g_loss, None


class L1_plus_perceptualLoss(nn.Module):
    def __init__(self, lambda_L1, lambda_perceptual, perceptual_layers, gpu_ids, percep_is_l1):
        super().__init__()
        self.lambda_L1 = lambda_L1
        self.lambda_perceptual = lambda_perceptual
        self.gpu_ids = gpu_ids
        self.percep_is_l1 = percep_is_l1
        self.perceptual_layers = perceptual_layers
        self.vgg_submodel = _vgg_slice(perceptual_layers)
        for p in self.vgg_submodel.parameters():
            p.requires_grad_(False)       # never optimised by the reference either
        self._eng = {}

    def vgg_engine(self, B, H, W):
        ops = runtime.get_ops(self.vgg_submodel[0].weight.device)
        key = (B, H, W, str(ops.device))
        if key not in self._eng:
            self._eng.clear()
            v, two = self.vgg_submodel, self.perceptual_layers >= 2
            self._eng[key] = VggEngine(ops, v[0].weight, v[0].bias, v[2].weight if two else None,
                                       v[2].bias if two else None, B, H, W, layers=self.perceptual_layers)
        return self._eng[key]

    def forward(self, inputs, targets):
        if self.lambda_L1 == 0 and self.lambda_perceptual == 0:
            z = torch.zeros(1, device=inputs.device)
            return z, torch.zeros(1), torch.zeros(1)
        return _LossFn.apply(self, inputs, targets)
