"""Batched pose-augmentation generation, the loop body of the reference's ``aug.py`` (:42-71) on the device.

``aug.py`` runs the generator at batch 1, copies the float image to the host, de-normalises and colour-swaps it with
numpy / cv2 and lets ``cv2.imwrite`` round and encode it, one image at a time with a device synchronisation each. Here:

  * a whole batch goes through the generator once; inputs may be the compact form of mmhand_b200.loader (uint8 frames,
    keypoints -- turned into the six tensors on the device);
  * ``images_to_bgr8`` produces the bytes ``cv2.imwrite`` would store -- ``saturate_cast<uchar>(cvRound((x * 0.5 + 0.5) *
    255))``, BGR, HWC -- on the device (mmh_image_pack_bgr8), so that only uint8 images cross the bus (4x fewer bytes);
  * ``ImageWriter`` copies them into a ring of pinned host buffers asynchronously and hands the PNG encoding + file
    write (cv2.imwrite releases the GIL) to a pool of host threads, so that the GPU never waits for the disk.
"""
import os
import queue
from concurrent.futures import ThreadPoolExecutor

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


@torch.no_grad()
def generate_batch(model, batch, host_out=None):
    """``generate`` on a batch dict of the loader: the reference's fp32 tensors, or the compact form
    (mmhand_b200.loader.CompactBatch materialises 'H1', 'P1', ... on the device on access)."""
    dev = next(model.parameters()).device
    t = lambda k: batch[k].to(dev, non_blocking=True).float()
    return generate(model, t('H1'), t('P1'), t('P2'), t('D1'), t('D2'), host_out=host_out)


class ImageWriter:
    """Asynchronous write-out of generated batches: ``submit(images_bgr8_on_device, paths)`` returns at once.

    The device-to-host copy goes to one of ``depth`` pinned buffers on the submitting stream; a host thread of the pool
    waits for that copy only (a CUDA event), then encodes and writes the files with ``cv2.imwrite`` -- the same call, on
    the same bytes, as aug.py:71. ``close()`` (or leaving the ``with`` block) waits for everything and re-raises the
    first write error."""

    def __init__(self, workers=4, depth=3):
        import cv2
        self.cv2 = cv2
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self.depth, self.slots = depth, [None] * depth
        self.free = queue.Queue()                # indices of the pinned buffers no writer thread is reading
        for i in range(depth):
            self.free.put(i)
        self.pending = []

    def _slot(self, shape, cuda):
        i = self.free.get()                      # back-pressure: at most `depth` batches in flight; a buffer is handed
        buf = self.slots[i]                      # out again only after ITS writer released it (completion order is
        if buf is None or tuple(buf.shape) != tuple(shape):          # not submission order with several workers)
            buf = torch.empty(shape, dtype=torch.uint8)
            if cuda:
                buf = buf.pin_memory()
            self.slots[i] = buf
        return i, buf

    def submit(self, images_bgr8, paths):
        assert images_bgr8.dtype == torch.uint8 and images_bgr8.dim() == 4 and len(paths) == images_bgr8.shape[0]
        cuda = images_bgr8.is_cuda
        slot, buf = self._slot(images_bgr8.shape, cuda)
        buf.copy_(images_bgr8, non_blocking=True)
        ev = None
        if cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(images_bgr8.device))
        self.pending.append(self.pool.submit(self._write, slot, buf, ev, list(paths)))

    def _write(self, slot, buf, ev, paths):
        try:
            if ev is not None:
                ev.synchronize()
            arr = buf.numpy()
            for i, p in enumerate(paths):
                d = os.path.dirname(p)
                if d:
                    os.makedirs(d, exist_ok=True)
                if not self.cv2.imwrite(p, arr[i]):
                    raise IOError("cv2.imwrite failed for %s" % p)
        finally:
            self.free.put(slot)

    def close(self):
        errs = []
        for f in self.pending:
            try:
                f.result()
            except Exception as e:      # keep draining: every slot must be released
                errs.append(e)
        self.pending = []
        self.pool.shutdown(wait=True)
        if errs:
            raise errs[0]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
