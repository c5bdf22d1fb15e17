"""Batched pose-augmentation generation, the loop body of the reference's ``aug.py`` (:42-71) on the device.

``aug.py`` runs the generator at batch 1, copies the float image to the host, de-normalises and colour-swaps it
with numpy / cv2 and lets ``cv2.imwrite`` round it. Here a batch goes through the generator once and
``images_to_bgr8`` produces the bytes ``cv2.imwrite`` would store -- ``saturate_cast<uchar>(cvRound((x * 0.5 + 0.5) *
255))``, BGR, HWC -- on the device (mmh_image_pack_bgr8), so that only uint8 images cross the bus.
"""
import torch

from . import runtime


def images_to_bgr8(images, out=None):
    """images: fp32 [B, 3, H, W] RGB in (-1, 1) on the GPU -> uint8 [B, H, W, 3] BGR (same device)."""
    assert images.dim() == 4 and images.shape[1] == 3
    ops = runtime.get_ops(images.device if images.is_cuda else None)
    src = images.to(ops.device, torch.float32).contiguous()
    B, _, H, W = src.shape
    if out is None:
        out = torch.empty(B, H, W, 3, dtype=torch.uint8, device=ops.device)
    ops.image_pack_bgr8(src, out)
    return out


@torch.no_grad()
def generate(model, H1, P1, P2, D1, D2, host_out=None):
    """One aug.py iteration for a whole batch: ``model([H1, cat(P1, P2), cat(D1, D2)])`` in eval mode, then the BGR
    uint8 images; with ``host_out`` (pinned uint8 [B, H, W, 3]) the result is copied to the host asynchronously."""
    fake = model([H1, torch.cat((P1, P2), 1), torch.cat((D1, D2), 1)])
    img = images_to_bgr8(fake)
    if host_out is not None:
        host_out.copy_(img, non_blocking=True)
        return host_out
    return img
