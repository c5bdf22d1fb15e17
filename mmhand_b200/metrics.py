"""Evaluator hooks of the reference's benchmark harness on the device.

``ssim(img1, img2, window_size=11, size_average=True)`` has the signature and semantics of
``pytorch_ssim.ssim`` (baselines/quantitative_on_benchmarks/pytorch_ssim/__init__.py:65-73), which
``Evaluator._get_SSIM_score`` calls for every generated image (utils.py:100-111); one fused kernel (mmh_ssim) replaces
its five depthwise convolutions and the map arithmetic."""
import torch

from . import runtime


def ssim(img1, img2, window_size=11, size_average=True):
    assert img1.shape == img2.shape and img1.dim() == 4
    ops = runtime.get_ops(img1.device if img1.is_cuda else None)
    a = img1.to(ops.device, torch.float32).contiguous()
    b = img2.to(ops.device, torch.float32).contiguous()
    B, C, H, W = a.shape
    if size_average:
        acc = torch.zeros(1, dtype=torch.float32, device=ops.device)
        ops.ssim(a, b, window_size, 1.5, acc, None)
        return (acc / float(B * C * H * W)).reshape(())
    per = torch.zeros(B, dtype=torch.float32, device=ops.device)
    ops.ssim(a, b, window_size, 1.5, None, per)
    return per
